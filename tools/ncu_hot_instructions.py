"""Hottest stalled SASS instructions of one kernel of an ncu report (needs --import-source on / -lineinfo), and for
consumers stalled on the long scoreboard the distance back to the load that feeds them -- how the sinking of the gather
loads in smooth3d16_kernel was found.
usage: python tools/ncu_hot_instructions.py report.ncu-rep <kernel id> [top=30]"""
import csv
import re
import subprocess
import sys

rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", ":::" + kid], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
print(rows[0][1][:160])
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
data = [r for r in rows[2:] if len(r) == len(hdr) and r[0].startswith("0x")]
tot = sum(int(r[ci["# Samples"]]) for r in data)
keys = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("instructions %d, samples %d" % (len(data), tot))
hot = sorted(range(len(data)), key=lambda i: -int(data[i][ci["# Samples"]]))[:top]
for i in sorted(hot):
    r = data[i]
    src = r[ci["Source"]].strip()
    st = sorted(((int(r[ci[k]] or 0), k.replace("stall_", "")) for k in keys), reverse=True)[:2]
    note = ""
    if st and st[0][1] == "long_sb":
        regs = re.findall(r"R(\d+)", src)[1:]
        for j in range(i - 1, max(i - 600, -1), -1):
            s2 = data[j][ci["Source"]].strip()
            if s2.startswith("LDG") and any(re.search(r"LDG\S* R%s," % g, s2) for g in regs):
                note = "  <- load %d instructions earlier" % (i - j)
                break
    print("%5d %5.1f%% %-58s %s%s" % (i, 100.0 * int(r[ci["# Samples"]]) / tot, src[:58], st, note))

#!/bin/bash
# ncu --set full of ONE whole V-cycle (the 4th: after the 3 warm-up cycles) of a bench.py config whose hierarchy has L levels,
# summarised by tools/ncu_summary.py.  usage: tools/ncu_cycle.sh CONFIG L OUT.txt   (3L - 2 smoother / face-residual launches per cycle)
set -u
CFG=$1; L=$2; OUT=$3
N=$((3 * L - 2))
rm -f /tmp/ncu_cycle_$CFG.ncu-rep
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'smooth|face_residual' --launch-skip $((3 * N)) -c $N -f -o /tmp/ncu_cycle_$CFG \
  python bench.py --config $CFG --cycle-only --no-cpu-baseline --no-other-configs --steps 1 --warmup 3 > /tmp/ncu_cycle_$CFG.log 2>&1
python tools/ncu_summary.py /tmp/ncu_cycle_$CFG.ncu-rep > $OUT 2>&1
grep -c "^====" $OUT

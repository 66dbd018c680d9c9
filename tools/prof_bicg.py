import os, sys, time
sys.path.insert(0, "/root/repo")
import pressurepoissonsolver_b200 as pps
ctx = pps.Context(0)
mesh = pps.Mesh.load("/root/repo/tests/golden/meshes/4uni.bin", 3).refine_leaves(1)
h = pps.Hierarchy.from_mesh(ctx, mesh, 16)
f, x = h.new_vec(0), h.new_vec(0)
h.init_trig_rhs(f)
opts = pps.CycleOpts.default()
h.bicgstab(f, x, opts, tol=1e-10, max_it=100)
for rep in range(2):
    x.set(0.0); ctx.sync(); t0 = time.perf_counter()
    its, rel = h.bicgstab(f, x, opts, tol=1e-10, max_it=100)
    ctx.sync(); print("bicgstab", its, rel, (time.perf_counter() - t0) * 1e3, "ms")
x.set(0.0)
ctx.profile_begin()
its, rel = h.bicgstab(f, x, opts, tol=1e-10, max_it=100)
prof = ctx.profile_end()
agg = {}
for name, lvl, ms in prof:
    a = agg.setdefault((name, lvl), [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(v[1] for v in agg.values())
print("total kernel ms", tot)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(k, v[0], round(v[1], 4), round(v[1] / v[0] * 1e3, 1), "us each")

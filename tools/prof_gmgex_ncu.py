"""One V-cycle of the reference's Neumann GMG example (16^3 patches, multi_refine_8 --divide 2) for an ncu launch list:
   ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/prof_gmgex_ncu.py"""
import os
import sys
sys.path.insert(0, os.getcwd())
import pressurepoissonsolver_b200 as pps
ctx = pps.Context(0)
mesh = pps.Mesh.load("tests/golden/meshes/3d_multi_refine_8.bin", 3).refine_leaves(2)
mesh.set_neumann(True)
h = pps.Hierarchy.from_mesh(ctx, mesh, 16)
f, u = h.new_vec(0), h.new_vec(0)
h.init_neumann_rhs(f, None, "gauss")
i, v = h.integrate(f)
f.shift(-i / v)
o = pps.CycleOpts.default(use_graph=0)
for _ in range(int(os.environ.get("CYCLES", "2"))):
    h.vcycle(f, u, o)
ctx.sync()
print("patches per level", [h.npatch(l) for l in range(h.nlevels)])

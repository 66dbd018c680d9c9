import sys, os, json
sys.path.insert(0, os.getcwd())
import pressurepoissonsolver_b200 as pps
ctx = pps.Context(0)
for neumann in (False, True):
    mesh = pps.Mesh.load("tests/golden/meshes/3d_multi_refine_8.bin", 3).refine_leaves(2)
    if neumann: mesh.set_neumann(True)
    h = pps.Hierarchy.from_mesh(ctx, mesh, 16)
    f, u = h.new_vec(0), h.new_vec(0)
    if neumann:
        h.init_neumann_rhs(f, None, "gauss"); i, v = h.integrate(f); f.shift(-i / v)
    else:
        h.init_trig_rhs(f)
    for _ in range(3): h.vcycle(f, u)
    ctx.sync(); ctx.timer_start()
    for _ in range(10): h.vcycle(f, u)
    ms = ctx.timer_stop() / 10
    ctx.profile_begin()
    for _ in range(5): h.vcycle(f, u)
    agg = {}
    for name, lvl, k in ctx.profile_end():
        a = agg.setdefault((name, lvl), 0.0); agg[(name, lvl)] = a + k / 5
    print("neumann", neumann, "ms/cycle", ms, "cells", h.ncells(0), [h.npatch(l) for l in range(h.nlevels)])
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:12]: print("   ", k, round(v, 4))
    h.close(); mesh.close()

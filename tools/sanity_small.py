"""Small end-to-end exercise of every specialised kernel (16^3, 32^3 cluster, 32^3 Neumann, 2D 32^2, lean face residual,
multi-sweep variants, BiCGStab with emitted faces) against the oracle: meant to be run under compute-sanitizer
(memcheck / racecheck), where the full test-suite would take too long."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gmg_oracle as go  # noqa: E402
import pressurepoissonsolver_b200 as pps  # noqa: E402

MESHES = os.path.join(ROOT, "tests", "golden", "meshes")
ctx = pps.Context(0)
cases = [("2refine.bin", 3, 16, 1, False), ("2uni.bin", 3, 32, 0, False), ("2d2ref.bin", 2, 32, 1, False), ("2uni.bin", 3, 32, 0, True),
         ("2refine.bin", 3, 16, 0, True)]  # the last: Neumann instantiation of smooth3d16_kernel on a refined mesh
REPS = int(os.environ.get("REPS", "30"))
if len(sys.argv) > 1:
    cases = [cases[int(a)] for a in sys.argv[1:]]
for mesh_file, D, n, divide, neumann in cases:
    mesh = pps.Mesh.load(os.path.join(MESHES, mesh_file), D)
    if neumann:
        mesh.set_neumann(True)
    mesh.refine_leaves(divide)
    h = pps.Hierarchy.from_mesh(ctx, mesh, n)
    levels = go.build_hierarchy(os.path.join(MESHES, mesh_file), D, n, divide, neumann=neumann)
    fn = np.random.default_rng(5).standard_normal(levels[0].shape)
    if neumann:
        fn -= fn.mean()
    f, u = h.new_vec(0, fn), h.new_vec(0)
    h.vcycle(f, u, pps.CycleOpts.default(use_graph=0))
    e1 = np.linalg.norm(u.download() - go.vcycle(levels, fn).ravel()) / np.linalg.norm(u.download())
    h.vcycle(f, u, pps.CycleOpts.default(pre_sweeps=2, post_sweeps=2, use_graph=0))
    e2 = np.linalg.norm(u.download() - go.vcycle(levels, fn, pre=2, post=2).ravel()) / np.linalg.norm(u.download())
    # run-to-run determinism (a shared-memory or halo race would show up as differing bits), graph replay included
    ref_bits = None
    for rep in range(REPS):
        h.vcycle(f, u, pps.CycleOpts.default(use_graph=rep & 1))
        bits = u.download().tobytes()
        ref_bits = ref_bits or bits
        assert bits == ref_bits, "V-cycle result changed between runs (rep %d)" % rep
    its, rel = h.bicgstab(f, u, tol=1e-8, max_it=30)
    print("%-14s D=%d n=%d neumann=%d: V(1,1) %.1e  V(2,2) %.1e  bicgstab %d its %.1e" % (mesh_file, D, n, neumann, e1, e2, its, rel), flush=True)
    assert e1 < 1e-10 and e2 < 1e-10
    h.close()
    mesh.close()
print("ok")

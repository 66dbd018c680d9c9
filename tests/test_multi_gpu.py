"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): one process per GPU, patches
partitioned along the Morton curve, NCCL halo exchange.  The gathered distributed V-cycle / BiCGStab
results must equal the reference golden vectors and the single-GPU path."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import MESHES, ROOT, load_golden, rel_l2

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _worker(rank, world, idfile, mesh_file, D, n, divide, f_global, ret, env=None):
    os.environ.update(env or {})  # exchange-path switches are read when the library first needs them
    sys.path.insert(0, ROOT)
    import time
    import pressurepoissonsolver_b200 as pps
    ctx = pps.Context(rank)
    if rank == 0:
        uid = pps.comm_unique_id()
        with open(idfile + ".tmp", "wb") as fh:
            fh.write(uid)
        os.rename(idfile + ".tmp", idfile)
    else:
        while not os.path.exists(idfile):
            time.sleep(0.05)
        uid = open(idfile, "rb").read()
    ctx.comm_init(uid, rank, world)
    mesh = pps.Mesh.load(os.path.join(MESHES, mesh_file), D)
    if (env or {}).get("TEST_NEUMANN"):
        mesh.set_neumann(True)
    mesh.refine_leaves(divide)
    part = pps.Partition(mesh, n, rank, world, min_patches_per_rank=2)
    h = pps.Hierarchy.from_partition(ctx, part)
    pl = part.level(0)
    nc = n ** D
    fg = np.asarray(f_global).reshape(-1, nc)
    f, u, r = h.new_vec(0, fg[pl["owned_global"]]), h.new_vec(0), h.new_vec(0)
    out = {"owned": pl["owned_global"], "ndist": part.ndist}
    h.vcycle(f, u)
    out["vcycle"] = u.download().reshape(-1, nc)
    g = pps.CycleOpts.default(use_graph=2)  # NCCL exchanges captured in the CUDA graph
    ug = h.new_vec(0)
    h.vcycle(f, ug, g)
    h.vcycle(f, ug, g)
    out["vcycle_graph"] = ug.download().reshape(-1, nc)
    # several sweeps per visit: the in-kernel hand-over of a level's first exchanges mixed with the separate push / wait
    # kernels of the later ones, all on the same generation counters
    h.vcycle(f, ug, pps.CycleOpts.default(use_graph=2, pre_sweeps=2, post_sweeps=2, coarse_sweeps=2))
    out["vcycle22"] = ug.download().reshape(-1, nc)
    out["fnorm"] = f.two_norm()
    out["integral"] = h.integrate(f)  # (Domain::integrate(f), Domain::volume()) summed over the ranks
    h.apply(0, f, r)
    out["apply"] = r.download().reshape(-1, nc)
    x = h.new_vec(0)
    its, rel = h.bicgstab(f, x, tol=1e-12, max_it=100)
    out["its"], out["x"] = its, x.download().reshape(-1, nc)
    ret[rank] = out
    h.close()
    ctx.close()


def run_distributed(world, mesh_file, D, n, divide, f_global, env=None):
    import tempfile
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    ret = mpc.Manager().dict()
    with tempfile.TemporaryDirectory() as tmp:
        idfile = os.path.join(tmp, "nccl_id")
        procs = [mpc.Process(target=_worker, args=(r, world, idfile, mesh_file, D, n, divide, f_global, ret, env)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
            assert p.exitcode == 0
    nc = n ** D
    P = sum(len(ret[r]["owned"]) for r in range(world))
    gather = lambda key: np.concatenate([ret[r][key] for r in range(world)])[np.argsort(np.concatenate([ret[r]["owned"] for r in range(world)]))]  # noqa: E731
    return ret, gather, P * nc


@pytest.mark.parametrize("name", ["3d_2refine_d1_n4", "3d_multi_refine_n4", "2d_multi_refine_8_n4", "3d_2refine_n8"])
def test_distributed_cycle_matches_reference(name):
    world = min(_ngpu(), 4)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    g = load_golden(name)
    ret, gather, ncells = run_distributed(world, str(g["mesh"]), int(g["D"]), int(g["n"]), int(g["divide"]), g["rhs_f"])
    assert ret[0]["ndist"] >= 1
    assert ncells == g["rhs_f"].size
    assert rel_l2(gather("vcycle"), g["vcycle"]) < 1e-12
    assert np.array_equal(gather("vcycle_graph"), gather("vcycle"))
    assert abs(ret[0]["fnorm"] / np.linalg.norm(g["rhs_f"]) - 1) < 1e-13
    # unit square / cube; every rank holds the same global sums
    assert abs(ret[0]["integral"][1] - 1.0) < 1e-12 and all(ret[r]["integral"] == ret[0]["integral"] for r in range(world))
    assert ret[0]["its"] == int(g["bicgstab_info"][0])
    assert rel_l2(gather("x"), g["bicgstab_u"]) < 1e-10


@pytest.mark.parametrize("mesh_file,D,n,divide", [("2refine.bin", 3, 16, 1), ("2refine.bin", 3, 32, 0), ("2d_multi_refine_8.bin", 2, 32, 0)])
def test_distributed_cycle_specialised_kernels_match_oracle(mesh_file, D, n, divide):
    """The kernels specialised for 16^3 (smooth3d16), 32^3 (cluster-pair smooth3d32c) and 2D 32^2 (smooth2d32) patches with the
    in-kernel halo hand-over (one launch over interior + boundary patches, CTAs / warps poll the peers' flags themselves),
    against the oracle on refined meshes."""
    world = min(_ngpu(), 2)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gmg_oracle as go
    levels = go.build_hierarchy(os.path.join(MESHES, mesh_file), D, n, divide)
    fn = np.random.default_rng(11).standard_normal(levels[0].shape)
    ret, gather, ncells = run_distributed(world, mesh_file, D, n, divide, fn)
    assert ret[0]["ndist"] >= 1 and ncells == fn.size
    ref = go.vcycle(levels, fn)
    assert rel_l2(gather("vcycle"), ref.reshape(-1, n ** D)) < 1e-12
    assert np.array_equal(gather("vcycle_graph"), gather("vcycle"))
    ref22 = go.vcycle(levels, fn, pre=2, post=2, coarse_sweeps=2)
    assert rel_l2(gather("vcycle22"), ref22.reshape(-1, n ** D)) < 1e-12


@pytest.mark.parametrize("env", [{"TGPU_PUSH_IN_KERNEL": "0"}, {"TGPU_HALO_IN_KERNEL": "0"}, {"TGPU_P2P": "0"}])
def test_alternative_exchange_paths_match_reference(env):
    """the exchange paths behind the diagnostic switches - a separate push launch instead of the consumer launch's own
    prologue, separate wait / signal kernels with split interior / boundary launches, and pack -> ncclSend/ncclRecv -> unpack
    (the fallback when peer mapping is unavailable) - give the same cycle"""
    world = min(_ngpu(), 2)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gmg_oracle as go
    levels = go.build_hierarchy(os.path.join(MESHES, "2refine.bin"), 3, 16, 1)
    fn = np.random.default_rng(12).standard_normal(levels[0].shape)
    ret, gather, ncells = run_distributed(world, "2refine.bin", 3, 16, 1, fn, env=env)
    assert rel_l2(gather("vcycle"), go.vcycle(levels, fn).reshape(-1, 16 ** 3)) < 1e-12
    assert np.array_equal(gather("vcycle_graph"), gather("vcycle"))


def test_distributed_neumann_cycle_matches_oracle():
    """Neumann domain boundaries across GPUs: mixed levels (the plain and the Neumann instantiation of smooth3d16_kernel on
    the interior / boundary ranges of every rank, separate hand-over kernels) against the oracle"""
    world = min(_ngpu(), 2)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gmg_oracle as go
    levels = go.build_hierarchy(os.path.join(MESHES, "3uni.bin"), 3, 16, 1, neumann=True)
    fn = np.random.default_rng(13).standard_normal(levels[0].shape)
    fn -= fn.mean()
    ret, gather, ncells = run_distributed(world, "3uni.bin", 3, 16, 1, fn, env={"TEST_NEUMANN": "1"})
    assert ret[0]["ndist"] >= 1 and ncells == fn.size
    assert rel_l2(gather("vcycle"), go.vcycle(levels, fn).reshape(-1, 16 ** 3)) < 1e-10
    assert np.array_equal(gather("vcycle_graph"), gather("vcycle"))
    assert rel_l2(gather("vcycle22"), go.vcycle(levels, fn, pre=2, post=2, coarse_sweeps=2).reshape(-1, 16 ** 3)) < 1e-10


def _replicated_worker(rank, world, idfile, ret):
    sys.path.insert(0, ROOT)
    import time
    import pressurepoissonsolver_b200 as pps
    ctx = pps.Context(rank)
    if rank == 0:
        uid = pps.comm_unique_id()
        with open(idfile + ".tmp", "wb") as fh:
            fh.write(uid)
        os.rename(idfile + ".tmp", idfile)
    else:
        while not os.path.exists(idfile):
            time.sleep(0.05)
        uid = open(idfile, "rb").read()
    ctx.comm_init(uid, rank, world)
    mesh = pps.Mesh.load(os.path.join(MESHES, "2uni.bin"), 3)
    part = pps.Partition(mesh, 8, rank, world, min_patches_per_rank=64)  # 8 patches < world * 64: every level replicated
    h = pps.Hierarchy.from_partition(ctx, part)
    f, u = h.new_vec(0), h.new_vec(0)
    h.init_trig_rhs(f)
    h.vcycle(f, u)
    ret[rank] = {"ndist": part.ndist, "integral": h.integrate(f), "fnorm": f.two_norm(), "u": u.download()}
    h.close()
    ctx.close()


def test_replicated_hierarchy_sums_are_not_multiplied_by_the_rank_count():
    """a mesh too small to distribute (ndist == 0): every rank owns every patch, so Domain::integrate / volume and the
    norms must come out once, not nranks times (no all-reduce on replicated levels)"""
    world = min(_ngpu(), 2)
    if world < 2:
        pytest.skip("needs >= 2 GPUs")
    import tempfile
    import torch.multiprocessing as mp
    mpc = mp.get_context("spawn")
    ret = mpc.Manager().dict()
    with tempfile.TemporaryDirectory() as tmp:
        procs = [mpc.Process(target=_replicated_worker, args=(r, world, os.path.join(tmp, "id"), ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(300)
            assert p.exitcode == 0
    g = load_golden("3d_2uni_n8")
    for r in range(world):
        assert ret[r]["ndist"] == 0
        assert abs(ret[r]["integral"][1] - 1.0) < 1e-13  # volume of the unit cube
        assert abs(ret[r]["fnorm"] / np.linalg.norm(g["rhs_f"]) - 1) < 1e-13
        assert rel_l2(ret[r]["u"], g["vcycle"]) < 1e-12

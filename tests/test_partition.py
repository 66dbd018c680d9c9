"""Host-side multi-GPU logic (CPU): Morton partition, local neighbour tables with halo slots, and the
halo-exchange plan.  The plan is executed here with numpy (all ranks in one process) and with
torch.distributed/gloo (world_size 2, two real processes); the per-rank operator results, computed by
the oracle on the LOCAL tables with exchanged halo faces, must equal the global operator."""
import os
import socket

import numpy as np
import pytest

import gmg_oracle as go
import pressurepoissonsolver_b200 as pps
from conftest import MESHES

CASES = [("2refine.bin", 3, 4, 2), ("3uni.bin", 3, 4, 0), ("2d_multi_refine_8.bin", 2, 4, 1), ("2d2ref.bin", 2, 8, 2)]


def local_level(pl, D, n):
    L = go.Level(D, n, pl["npatch"])
    for k in ("nbr_type", "nbr_idx", "orth_on_coarse", "spacings", "starts", "neumann", "parent_idx", "orth_on_parent"):
        getattr(L, k)[...] = pl[k]
    return L


def pack(pl, peer, u_loc, D):
    return [np.array(go.face(u_loc[p:p + 1], D, int(s))[0]) for p, s in zip(peer["send_patch"], peer["send_side"])]


def unpack(peer, faces, u_loc, D):
    for slot, s, fa in zip(peer["recv_slot"], peer["recv_side"], faces):
        go.face(u_loc[slot:slot + 1], D, int(s))[0][...] = fa


@pytest.mark.parametrize("mesh_file,D,n,divide", CASES)
@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_partition_and_halo_plan(mesh_file, D, n, divide, nranks):
    path = os.path.join(MESHES, mesh_file)
    mesh = pps.Mesh.load(path, D).refine_leaves(divide)
    glob = go.build_hierarchy(path, D, n, divide)
    parts = [pps.Partition(mesh, n, r, nranks, min_patches_per_rank=2) for r in range(nranks)]
    ndist = parts[0].ndist
    assert ndist >= 1 and all(p.ndist == ndist and p.nlevels == len(glob) for p in parts)
    rng = np.random.default_rng(5)
    for l, G in enumerate(glob):
        pls = [p.level(l) for p in parts]
        if l >= ndist:  # replicated: the global level on every rank
            for pl in pls:
                assert pl["n_owned"] == G.P and pl["n_halo"] == 0 and np.array_equal(pl["nbr_idx"], G.nbr_idx)
            continue
        owned = np.concatenate([pl["owned_global"] for pl in pls])
        assert sorted(owned.tolist()) == list(range(G.P))  # every patch owned exactly once
        counts = [pl["n_owned"] for pl in pls]
        if l == 0 and mesh_file == "3uni.bin" and nranks in (2, 8):
            assert max(counts) == min(counts)  # uniform mesh: perfectly balanced
        u = rng.standard_normal(G.shape)
        f = rng.standard_normal(G.shape)
        Au, Su = go.apply_op(G, u), go.smooth(G, f, u)
        locs = []
        for pl in pls:
            ul = np.zeros((pl["npatch"],) + G.shape[1:])
            ul[:pl["n_owned"]] = u[pl["owned_global"]]
            locs.append(ul)
        # execute the plan: rank r's k-th send list to peer q must match q's receive list from r
        for r, pl in enumerate(pls):
            for peer in pl["peers"]:
                q = peer["peer"]
                back = [x for x in pls[q]["peers"] if x["peer"] == r][0]
                assert len(peer["send_patch"]) == len(back["recv_slot"])
                assert np.array_equal(pl["owned_global"][peer["send_patch"]],
                                      pls[q]["halo_global"][back["recv_slot"] - pls[q]["n_owned"]])
                assert np.array_equal(peer["send_side"], back["recv_side"])
                unpack(back, pack(pl, peer, locs[r], D), locs[q], D)
        for r, pl in enumerate(pls):
            LL = local_level(pl, D, n)
            no = pl["n_owned"]
            fl = np.zeros_like(locs[r])
            fl[:no] = f[pl["owned_global"]]
            assert np.array_equal(go.apply_op(LL, locs[r])[:no], Au[pl["owned_global"]])
            assert np.allclose(go.smooth(LL, fl, locs[r])[:no], Su[pl["owned_global"]], rtol=0, atol=1e-13)
            # parents of owned patches are local (restriction / prolongation need no communication)
            if l + 1 < len(glob):
                nxt = parts[r].level(l + 1)
                gp = G.parent_idx[pl["owned_global"]]
                assert np.array_equal(nxt["owned_global"][pl["parent_idx"][:no]], gp)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _gloo_worker(rank, world, port, mesh_file, D, n, divide, ret):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    path = os.path.join(MESHES, mesh_file)
    mesh = pps.Mesh.load(path, D).refine_leaves(divide)
    glob = go.build_hierarchy(path, D, n, divide)
    part = pps.Partition(mesh, n, rank, world, min_patches_per_rank=2)
    G, pl = glob[0], part.level(0)
    rng = np.random.default_rng(11)  # same seed on both ranks: the global vector
    u = rng.standard_normal(G.shape)
    ul = np.zeros((pl["npatch"],) + G.shape[1:])
    ul[:pl["n_owned"]] = u[pl["owned_global"]]
    reqs, bufs = [], []
    for peer in pl["peers"]:  # grouped isend / irecv, like ncclGroupStart .. ncclGroupEnd on the GPU
        sb = torch.from_numpy(np.stack(pack(pl, peer, ul, D)))
        rb = torch.empty((len(peer["recv_slot"]),) + sb.shape[1:], dtype=torch.float64)
        reqs += [dist.isend(sb, peer["peer"]), dist.irecv(rb, peer["peer"])]
        bufs.append((peer, rb))
    for q in reqs:
        q.wait()
    for peer, rb in bufs:
        unpack(peer, list(rb.numpy()), ul, D)
    no = pl["n_owned"]
    ok = np.array_equal(go.apply_op(local_level(pl, D, n), ul)[:no], go.apply_op(G, u)[pl["owned_global"]])
    # norm of a distributed vector = local partial + all-reduce (Vector.h:283-297 MPI_Allreduce)
    t = torch.tensor([float(np.sum(ul[:no] ** 2))], dtype=torch.float64)
    dist.all_reduce(t)
    ok = ok and abs(t.item() - float(np.sum(u ** 2))) < 1e-9 * float(np.sum(u ** 2))
    ret[rank] = bool(ok)
    dist.destroy_process_group()


def test_halo_exchange_over_gloo_world_size_2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    port = _free_port()
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, "2refine.bin", 3, 4, 1, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert ret[0] and ret[1]


def test_bench_nvlink_block_counts_the_send_lists():
    """bench.py's `nvlink` figure = 8 bytes x face size x send-list length x 2 hand-overs per distributed level; on two ranks
    what one rank sends the other receives"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(MESHES), "..", "..", "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)

    class OneRankEnv:
        def reduce_max(self, *vals):
            return list(vals)

    mesh = pps.Mesh.load(os.path.join(MESHES, "2refine.bin"), 3).refine_leaves(1)
    parts = [pps.Partition(mesh, 8, r, 2, min_patches_per_rank=2) for r in range(2)]
    blocks = [bench.nvlink_block(OneRankEnv(), p, 3, 8, 1.0) for p in parts]
    for r, (p, b) in enumerate(zip(parts, blocks)):
        assert p.ndist >= 1 and len(b["faces_sent_per_level_rank0"]) == p.ndist
        sent = sum(len(q["send_patch"]) for l in range(p.ndist) for q in p.level(l)["peers"])
        recv_other = sum(len(q["recv_slot"]) for l in range(p.ndist) for q in parts[1 - r].level(l)["peers"])
        assert sent == recv_other > 0
        assert b["bytes_sent_per_rank_per_cycle"] == sent * 64 * 8 * 2
        assert abs(b["frac"] - b["achieved_gbs_per_rank"] / 900.0) < 1e-15
    for p in parts:
        p.close()
    mesh.close()

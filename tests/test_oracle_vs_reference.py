"""Pins oracle/gmg_oracle.py (the numpy restatement) against outputs of the reference's own
sources (tests/golden/*.npz, produced by tests/golden/make_golden.py through oracle/_ref/ref_gmg).
CPU only."""
import os

import numpy as np
import pytest

import gmg_oracle as go
from conftest import GOLDEN_CASES, MESHES, NEUMANN_CASES, load_golden, rel_l2

TOL = 1e-13  # restatement vs reference: only fp re-association differs


@pytest.fixture(scope="module", params=GOLDEN_CASES)
def case(request):
    g = load_golden(request.param)
    levels = go.build_hierarchy(os.path.join(MESHES, str(g["mesh"])), int(g["D"]), int(g["n"]), int(g["divide"]))
    return g, levels


def test_metadata_bit_exact(case):
    g, levels = case
    assert len(levels) == int(g["nlevels"])
    for l, L in enumerate(levels):
        for k in ("ids", "refine_level", "parent_id", "orth_on_parent", "parent_idx", "nbr_type", "nbr_ids",
                  "nbr_idx", "orth_on_coarse", "starts", "spacings"):
            assert np.array_equal(getattr(L, k), g["L%d_%s" % (l, k)]), (l, k)


def test_rhs(case):
    g, levels = case
    f, exact = go.trig_rhs(levels[0])
    assert rel_l2(f, g["rhs_f"]) < 1e-14
    assert rel_l2(exact, g["rhs_exact"]) < 1e-14


def test_level_operators(case):
    g, levels = case
    for l, L in enumerate(levels):
        u = g["L%d_in_u" % l].reshape(L.shape)
        f = g["L%d_in_f" % l].reshape(L.shape)
        assert rel_l2(go.apply_op(L, u), g["L%d_apply" % l]) < TOL
        assert rel_l2(go.smooth(L, f, u), g["L%d_smooth" % l]) < TOL
        # the reference's FFTW-planned solver and its first-party DFT solver agree
        assert rel_l2(g["L%d_smooth_fftw" % l], g["L%d_smooth" % l]) < TOL
        if l + 1 < len(levels):
            C = levels[l + 1]
            uc = g["L%d_in_uc" % l].reshape(C.shape)
            assert np.array_equal(go.restrict(L, C, u).ravel(), g["L%d_restrict" % l])
            assert np.array_equal(go.interpolate(L, C, uc, u).ravel(), g["L%d_interp" % l])


def test_vcycle_and_solvers(case):
    g, levels = case
    f = g["rhs_f"].reshape(levels[0].shape)
    assert rel_l2(go.vcycle(levels, f), g["vcycle"]) < 1e-12
    u, hist = go.vcycle_history(levels, f, 6)
    assert rel_l2(u, g["vhist_u"]) < 1e-12
    assert np.max(np.abs(hist / g["vhist"] - 1)) < 1e-9
    x, its = go.bicgstab(levels, f)
    assert its == int(g["bicgstab_info"][0])
    assert rel_l2(x, g["bicgstab_u"]) < 1e-10


@pytest.mark.parametrize("name", NEUMANN_CASES)
def test_neumann_boundaries(name):
    """Neumann domain sides (DCT-II/III, DCT-IV, DST-IV patch solves, zero mode of the all-Neumann
    coarsest patch): operator, smoother with both reference patch solvers, and a V-cycle."""
    g = load_golden(name)
    levels = go.build_hierarchy(os.path.join(MESHES, str(g["mesh"])), int(g["D"]), int(g["n"]), int(g["divide"]), neumann=True)
    assert len(levels) == int(g["nlevels"])
    for l, L in enumerate(levels):
        assert np.array_equal(L.neumann, g["L%d_neumann" % l])
        u = g["L%d_in_u" % l].reshape(L.shape)
        f = g["L%d_in_f" % l].reshape(L.shape)
        assert rel_l2(go.apply_op(L, u), g["L%d_apply" % l]) < TOL
        assert rel_l2(go.smooth(L, f, u), g["L%d_smooth" % l]) < 1e-12
        assert rel_l2(g["L%d_smooth_fftw" % l], g["L%d_smooth" % l]) < 1e-12
    assert rel_l2(go.vcycle(levels, g["rhs_f"].reshape(levels[0].shape)), g["vcycle"]) < 1e-11


def test_known_answers_from_reference_tests():
    """test/GMG.cpp:261-435 (disabled upstream): AvgRstr of a refined patch filled per-octant with
    the child's id gives octFill(child ids); un-refined patches copy 1:1; DrctIntp is the inverse
    pattern.  Mesh 2refine.bin, n = 8, as in the reference test."""
    levels = go.build_hierarchy(os.path.join(MESHES, "2refine.bin"), 3, 8, 0)
    fine, coarse = levels[0], levels[1]
    r = np.zeros(fine.shape)
    for p in range(fine.P):
        r[p] = fine.ids[p]
    rc = go.restrict(fine, coarse, r)
    for p in range(fine.P):
        c = fine.parent_idx[p]
        if fine.orth_on_parent[p] < 0:
            assert np.all(rc[c] == fine.ids[p])
        else:
            o = fine.orth_on_parent[p]
            blk = rc[c][(o >> 2 & 1) * 4:(o >> 2 & 1) * 4 + 4, (o >> 1 & 1) * 4:(o >> 1 & 1) * 4 + 4,
                        (o & 1) * 4:(o & 1) * 4 + 4]
            assert np.all(blk == fine.ids[p])
    uc = np.zeros(coarse.shape)
    for c in range(coarse.P):
        uc[c] = coarse.ids[c]
    uf = go.interpolate(fine, coarse, uc, np.zeros(fine.shape))
    for p in range(fine.P):
        assert np.all(uf[p] == fine.parent_id[p])


def test_second_order_convergence():
    """apps/3d/steady.cpp:536-560: the discretisation error of the converged solution vs the
    manufactured solution drops ~4x per uniform refinement (2uni -> 3uni at n = 4)."""
    errs = []
    for mesh in ("2uni.bin", "3uni.bin"):
        levels = go.build_hierarchy(os.path.join(MESHES, mesh), 3, 4, 0)
        f, exact = go.trig_rhs(levels[0])
        x, _ = go.bicgstab(levels, f)
        errs.append(np.max(np.abs(x - exact)))
    assert 3.0 < errs[0] / errs[1] < 5.0

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


CYCLE_GOLDENS = ["3d_2refine_n8", "2d_2d2ref_d1_n8", "3d_multi_refine_n4"]
# key in *_cycles.npz -> arguments of go.cycle (the "@" options of ref_gmg that produced it, tests/golden/make_golden.py)
CYCLE_VARIANT_ARGS = {
    "W": dict(cycle_type="W"), "W_p2m2c2": dict(cycle_type="W", pre=2, mid=2, post=1, coarse_sweeps=2),
    "V_max_levels2": dict(max_levels=2), "V_ppp2": dict(patches_per_proc=2.0), "W_max_levels2": dict(cycle_type="W", max_levels=2),
    "V_lambda": dict(lam=-3.5),
}


@pytest.mark.parametrize("name", CYCLE_GOLDENS)
def test_cycle_variants_vs_reference(name):
    """GMG::WCycle (GMG/WCycle.h:45-68), the level truncation of GMG::CycleFactory (GMG/CycleFactory3d.cpp:99-104) and the
    patch solver's lambda (DftPatchSolver.h:78,168), all against runs of the reference itself."""
    g = load_golden(name)
    c = load_golden(name + "_cycles")
    levels = go.build_hierarchy(os.path.join(MESHES, str(g["mesh"])), int(g["D"]), int(g["n"]), int(g["divide"]))
    f = g["rhs_f"].reshape(levels[0].shape)
    assert rel_l2(go.cycle(levels, f), g["vcycle"]) < 1e-12
    for key, kw in CYCLE_VARIANT_ARGS.items():
        assert rel_l2(go.cycle(levels, f, **kw), c["cycle_" + key]) < 1e-12, key
    # the truncated and W cycles really differ from the plain V cycle (the fixtures are not vacuous)
    assert rel_l2(c["cycle_W"], g["vcycle"]) > 1e-6 and rel_l2(c["cycle_V_max_levels2"], g["vcycle"]) > 1e-6
    lam = float(c["lambda"])
    for l, L in enumerate(levels):
        u, ff = c["L%d_in_u" % l].reshape(L.shape), c["L%d_in_f" % l].reshape(L.shape)
        assert rel_l2(go.smooth(L, ff, u, lam), c["L%d_smooth_lambda" % l]) < TOL
        assert rel_l2(c["L%d_smooth_lambda_fftw" % l], c["L%d_smooth_lambda" % l]) < TOL


def test_jacobi_is_a_sweep_on_the_assembled_diagonal():
    """jacobi(): the diagonal used must be the diagonal of the operator the reference applies (apply_op, pinned above):
    probe A with unit vectors on a small refined mesh, Dirichlet and Neumann."""
    for neumann in (False, True):
        levels = go.build_hierarchy(os.path.join(MESHES, "2d2ref.bin"), 2, 4, 0, neumann=neumann)
        L = levels[0]
        diag = np.zeros(L.shape)
        for idx in np.ndindex(*L.shape):
            e = np.zeros(L.shape)
            e[idx] = 1.0
            diag[idx] = go.apply_op(L, e)[idx]
        rng = np.random.default_rng(3)
        u, f = rng.standard_normal(L.shape), rng.standard_normal(L.shape)
        expect = u + 0.8 * (f - go.apply_op(L, u)) / diag
        assert rel_l2(go.jacobi(L, f, u, 0.8), expect) < 1e-14


@pytest.mark.parametrize("mesh,D,n,divide", [("2refine.bin", 3, 8, 0), ("3uni.bin", 3, 4, 0), ("2d2ref.bin", 2, 8, 1)])
def test_trilinear_interpolation_known_answer(mesh, D, n, divide):
    """test/GMG.cpp:465-600 (disabled upstream): the (tri)linear interpolator reproduces the linear field x + y/2 - z
    on uniform and refined two-level meshes; and its weights are the tables of GMG/TriLinIntp.cpp:110-190."""
    levels = go.build_hierarchy(os.path.join(MESHES, mesh), D, n, divide)
    fine, coarse = levels[0], levels[1]

    def field(L):
        out = np.zeros(L.shape)
        for p in range(L.P):
            c = [L.starts[p, a] + L.spacings[p, a] * (np.arange(n) + 0.5) for a in range(D)]
            if D == 3:
                out[p] = c[0][None, None, :] + 0.5 * c[1][None, :, None] - c[2][:, None, None]
            else:
                out[p] = c[0][None, :] + 0.5 * c[1][:, None]
        return out

    got = go.interpolate_trilinear(fine, coarse, field(coarse), np.zeros(fine.shape))
    assert rel_l2(got, field(fine)) < 1e-14
    if D == 3:
        # interior fine cell (1, 1, 1) of orthant 0 mixes the coarse cube (0..1)^3 with 27/9/9/3/9/3/3/1 over 64; the
        # corner cell (0, 0, 0) extrapolates with 125, -25, -25, 5, -25, 5, 5, -1 over 64
        p = int(np.nonzero(fine.orth_on_parent == 0)[0][0])
        uc = np.zeros(coarse.shape)
        cube = np.arange(1.0, 9.0).reshape(2, 2, 2)  # [z][y][x]
        uc[fine.parent_idx[p], :2, :2, :2] = cube
        uf = go.interpolate_trilinear(fine, coarse, uc, np.zeros(fine.shape))[p]
        w = np.array([27, 9, 9, 3, 9, 3, 3, 1]) / 64.0  # order x fastest: (0,0,0), (1,0,0), (0,1,0), ...
        assert abs(uf[1, 1, 1] - float(np.dot(w, cube.ravel()))) < 1e-14
        wc = np.array([125, -25, -25, 5, -25, 5, 5, -1]) / 64.0
        assert abs(uf[0, 0, 0] - float(np.dot(wc, cube.ravel()))) < 1e-14


def test_operator_against_the_references_assembled_matrix():
    """the reference carries a second, independent implementation of the 3D operator - the assembled matrix of
    MatrixHelper::formCRSMatrix with the per-side stencils of StencilHelper.h (coarse/fine weights included).  Golden
    3d_assembled_operator.npz holds that matrix applied to the per-level inputs of the 3D goldens (refined meshes, Dirichlet
    and Neumann); the reference's matrix-free operator (golden L*_apply) and the oracle must both reproduce it."""
    a = load_golden("3d_assembled_operator")
    seen = 0
    for name in GOLDEN_CASES + NEUMANN_CASES:
        g = load_golden(name)
        if int(g["D"]) != 3:
            continue
        neumann = name.endswith("_neumann")
        levels = go.build_hierarchy(os.path.join(MESHES, str(g["mesh"])), 3, int(g["n"]), int(g["divide"]), neumann=neumann)
        for l, L in enumerate(levels):
            m = a["%s_L%d_matapply" % (name, l)]
            assert rel_l2(g["L%d_apply" % l], m) < 1e-14, (name, l)
            assert rel_l2(go.apply_op(L, g["L%d_in_u" % l].reshape(L.shape)).ravel(), m) < 1e-14, (name, l)
            seen += 1
    assert seen == len(a.files) == 19

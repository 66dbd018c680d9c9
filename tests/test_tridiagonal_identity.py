"""The identity behind the patch solve of the GPU kernels: with D - 1 axes diagonalised by the reference's
DST-II / DST-III pair (DftPatchSolver.h:237-289, eigenvalues FftwPatchSolver.h:152-167), the remaining axis is a
tridiagonal system with the Dirichlet closure of StarPatchOp.h:46-64, and its two-sided elimination with one set of
tabulated multipliers (TriSolve in csrc/kernels.cuh, table built in csrc/tgpu.cu) gives the reference's patch solve.
Checked here in numpy against the oracle's patch solver; the GPU tests check the kernels themselves."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gmg_oracle as go  # noqa: E402


def tri_table(n, D):
    """[n/2 + 1][n^(D-1)] multipliers exactly as tgpu.cu builds them (long double there, float64 is enough here)"""
    H = n // 2
    lam = -4.0 * np.sin((np.arange(n) + 1) * np.pi / (2 * n)) ** 2
    mu = lam if D == 2 else (lam[None, :] + lam[:, None]).ravel()  # index k_a + n k_b
    tab = np.empty((H + 1, mu.size))
    a = np.zeros_like(mu)
    for j in range(H):
        a = 1.0 / (mu - (3.0 if j == 0 else 2.0) - (0.0 if j == 0 else a))
        tab[j] = a
    tab[H] = 1.0 / (1.0 - a * a)
    return tab


def tri_solve(v, tab, hs):
    """TriSolve::forward + ::backward on pencils v[n][npencil]"""
    n = v.shape[0]
    H = n // 2
    v = v.copy()
    sa = hs * tab[0]
    v[0] *= sa
    v[n - 1] *= sa
    for j in range(1, H):
        a = tab[j]
        v[j] = -a * v[j - 1] + v[j] * (hs * a)
        v[n - 1 - j] = -a * v[n - j] + v[n - 1 - j] * (hs * a)
    a, kap = tab[H - 1], tab[H]
    yt = kap * (-a * v[H] + v[H - 1])
    yb = kap * (-a * v[H - 1] + v[H])
    v[H - 1], v[H] = yt, yb
    for j in range(H - 2, -1, -1):
        v[j] = -tab[j] * v[j + 1] + v[j]
        v[n - 1 - j] = -tab[j] * v[n - 2 - j] + v[n - 1 - j]
    return v


def dst2_matrix(n):
    k, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    return np.sin(np.pi / n * (k + 1) * (j + 0.5))


def dst3_matrix(n):
    i, j = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    T = np.sin(np.pi / n * (i + 0.5) * (j + 1))
    T[:, n - 1] = 0.5 * (-1.0) ** np.arange(n)
    return T


@pytest.mark.parametrize("D,n", [(2, 8), (2, 32), (3, 8), (3, 16)])
def test_transforms_plus_tridiagonal_equals_reference_patch_solve(D, n):
    rng = np.random.default_rng(D * 100 + n)
    h = 1.0 / (4 * n)
    f = rng.standard_normal((n,) * D)
    F, Inv = dst2_matrix(n), dst3_matrix(n)
    # reference: transform every axis, divide by the eigenvalue sum, transform back, scale (2/n)^D
    lam = -4.0 / h**2 * np.sin((np.arange(n) + 1) * np.pi / (2 * n)) ** 2
    g = f.copy()
    for ax in range(D):
        g = np.moveaxis(np.tensordot(F, np.moveaxis(g, ax, 0), axes=1), 0, ax)
    ev = sum(lam.reshape([-1 if a == ax else 1 for a in range(D)]) for ax in range(D))
    g = g / ev
    for ax in range(D):
        g = np.moveaxis(np.tensordot(Inv, np.moveaxis(g, ax, 0), axes=1), 0, ax)
    ref = g * (2.0 / n) ** D
    # ours: transform D - 1 axes (array axes 1.., i.e. all but the first), eliminate along the first
    w = f.copy()
    for ax in range(1, D):
        w = np.moveaxis(np.tensordot(F, np.moveaxis(w, ax, 0), axes=1), 0, ax)
    tab = tri_table(n, D)
    # pencil index of the table: k_a + n k_b with a = last array axis (fastest), b = the one before
    w2 = w.reshape(n, -1)
    sol = tri_solve(w2, tab, h * h * (2.0 / n) ** (D - 1)).reshape(w.shape)
    for ax in range(1, D):
        sol = np.moveaxis(np.tensordot(Inv, np.moveaxis(sol, ax, 0), axes=1), 0, ax)
    assert np.linalg.norm(sol - ref) / np.linalg.norm(ref) < 1e-13


def test_matches_oracle_patch_solver_on_a_mesh():
    """the same through the oracle's own smoother on a single-patch level (zero neighbours: u = S^-1 f)"""
    mesh = os.path.join(ROOT, "tests", "golden", "meshes", "1uni.bin")
    if not os.path.exists(mesh):
        pytest.skip("1uni.bin fixture not present")
    n = 8
    levels = go.build_hierarchy(mesh, 3, n, 0)
    L = levels[-1]
    f = np.random.default_rng(3).standard_normal(L.shape)
    ref = go.smooth(L, f, np.zeros(L.shape)).reshape(n, n, n)  # [z][y][x]
    h = float(np.asarray(L.spacings).ravel()[0])
    F, Inv = dst2_matrix(n), dst3_matrix(n)
    w = f.reshape(n, n, n).copy()
    for ax in (1, 2):
        w = np.moveaxis(np.tensordot(F, np.moveaxis(w, ax, 0), axes=1), 0, ax)
    sol = tri_solve(w.reshape(n, -1), tri_table(n, 3), h * h * (2.0 / n) ** 2).reshape(n, n, n)
    for ax in (1, 2):
        sol = np.moveaxis(np.tensordot(Inv, np.moveaxis(sol, ax, 0), axes=1), 0, ax)
    assert np.linalg.norm(sol - ref) / np.linalg.norm(ref) < 1e-12

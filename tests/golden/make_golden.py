"""Generates tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (oracle/_ref/ref_gmg = the
reference's unmodified GMG sources compiled against single-rank shims, see oracle/Makefile)
on seeded inputs.  Run from the repo root in the build container (needs /root/reference only to
have built oracle/_ref):

    make -C oracle ref && python tests/golden/make_golden.py

The mesh files under tests/golden/meshes/ are byte copies of the reference's octree fixtures
(test/*.bin, apps/{2d,3d}/meshes/*.bin: data, not source) so that nothing reads /root/reference
at test time.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import gmg_oracle as go  # noqa: E402  (only used to read the reference's metadata dump)

REF = os.path.join(ROOT, "oracle", "_ref", "ref_gmg")

# name, D, mesh, divide, n
CASES = [
    ("3d_2uni_n8", 3, "2uni.bin", 0, 8),
    ("3d_2refine_n8", 3, "2refine.bin", 0, 8),
    ("3d_2refine_d1_n4", 3, "2refine.bin", 1, 4),
    ("3d_multi_refine_n4", 3, "multi_refine.bin", 0, 4),
    ("2d_2d2ref_d1_n8", 2, "2d2ref.bin", 1, 8),
    ("2d_multi_refine_8_n4", 2, "2d_multi_refine_8.bin", 0, 4),
]


def ref(D, mesh, div, n, *cmds, solver="dft"):
    subprocess.check_call([REF, str(D), os.path.join(HERE, "meshes", mesh), str(div), str(n), solver, *cmds])


# Neumann domain boundaries (ThundereggDomGen(..., neumann = true)): per-level operator and smoother with both
# patch solvers, and one V-cycle of a seeded right-hand side
NEUMANN_CASES = [
    ("3d_2refine_n8_neumann", 3, "2refine.bin", 0, 8),
    ("2d_2d2ref_d1_n8_neumann", 2, "2d2ref.bin", 1, 8),
    ("3d_2uni_n16_neumann", 3, "2uni.bin", 0, 16),
]


def neumann_cases():
    for name, D, mesh, div, n in NEUMANN_CASES:
        with tempfile.TemporaryDirectory() as tmp:
            t = lambda f: os.path.join(tmp, f)  # noqa: E731
            ref(D, mesh, div, n, "meta:" + t("meta"), solver="dft-neumann")
            levels = go.read_ref_meta(t("meta"))
            out = {"D": D, "n": n, "divide": div, "mesh": mesh, "nlevels": len(levels)}
            rng = np.random.default_rng(4321)
            for l, L in enumerate(levels):
                out["L%d_neumann" % l] = L.neumann
                u = rng.standard_normal(L.cells)
                f = rng.standard_normal(L.cells)
                u.tofile(t("u"))
                f.tofile(t("ff"))
                ref(D, mesh, div, n, "apply:%d:%s:%s" % (l, t("u"), t("au")),
                    "smooth:%d:%s:%s:%s" % (l, t("ff"), t("u"), t("su")), solver="dft-neumann")
                ref(D, mesh, div, n, "smooth:%d:%s:%s:%s" % (l, t("ff"), t("u"), t("su2")), solver="fftw-neumann")
                out["L%d_in_u" % l], out["L%d_in_f" % l] = u, f
                out["L%d_apply" % l] = np.fromfile(t("au"))
                out["L%d_smooth" % l] = np.fromfile(t("su"))
                out["L%d_smooth_fftw" % l] = np.fromfile(t("su2"))
            f = rng.standard_normal(levels[0].cells)
            f -= f.mean()  # compatible right-hand side (apps/3d/steady.cpp:330-334)
            f.tofile(t("f"))
            ref(D, mesh, div, n, "vcycle:%s:%s" % (t("f"), t("v")), solver="dft-neumann")
            out["rhs_f"] = f
            out["vcycle"] = np.fromfile(t("v"))
            np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
            print(name, [L.P for L in levels])


def neumann_init_case():
    """Init::initNeumann (apps/shared/Init.cpp:57-151) on the manufactured problems of apps/3d/steady.cpp:230-282 and the
    right-hand-side mean the app removes (apps/3d/steady.cpp:330-334), on the refined octree with 8^3 patches"""
    out = {"mesh": "2refine.bin", "D": 3, "n": 8, "divide": 0}
    with tempfile.TemporaryDirectory() as tmp:
        t = lambda k: os.path.join(tmp, k + ".bin")  # noqa: E731
        for prob in ("trig", "gauss"):
            txt = subprocess.check_output([REF, "3", os.path.join(HERE, "meshes", "2refine.bin"), "0", "8", "dft-neumann",
                                           "rhsn:%s:%s:%s" % (prob, t("f"), t("e"))], text=True)
            info = json.loads(txt.strip().splitlines()[-1])
            out["f_" + prob] = np.fromfile(t("f"))
            out["exact_" + prob] = np.fromfile(t("e"))
            out["fdiff_" + prob] = info["fdiff"]
            out["volume"] = info["volume"]
    np.savez_compressed(os.path.join(HERE, "3d_2refine_n8_neumann_init.npz"), **out)
    print("3d_2refine_n8_neumann_init", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})
    # 2D: Init::initNeumann2d, trig problem of apps/2d/steady.cpp:314-318, refined quadtree
    out = {"mesh": "2d2ref.bin", "D": 2, "n": 8, "divide": 1}
    with tempfile.TemporaryDirectory() as tmp:
        t = lambda k: os.path.join(tmp, k + ".bin")  # noqa: E731
        txt = subprocess.check_output([REF, "2", os.path.join(HERE, "meshes", "2d2ref.bin"), "1", "8", "dft-neumann",
                                       "rhsn:trig:%s:%s" % (t("f"), t("e"))], text=True)
        info = json.loads(txt.strip().splitlines()[-1])
        out["f_trig"], out["exact_trig"] = np.fromfile(t("f")), np.fromfile(t("e"))
        out["fdiff_trig"], out["volume"] = info["fdiff"], info["volume"]
    np.savez_compressed(os.path.join(HERE, "2d_2d2ref_d1_n8_neumann_init.npz"), **out)
    print("2d_2d2ref_d1_n8_neumann_init", {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


# Cycle variants and the patch solver's shift, from the reference's own WCycle / CycleFactory / DftPatchSolver:
# name -> "@" option string of ref_gmg (oracle/ref_driver.cpp)
CYCLE_VARIANTS = {
    "W": "W", "W_p2m2c2": "W,pre=2,mid=2,post=1,coarse=2", "V_max_levels2": "max_levels=2", "V_ppp2": "ppp=2",
    "W_max_levels2": "W,max_levels=2", "V_lambda": "lambda=-3.5",
}
CYCLE_CASES = [("3d_2refine_n8", 3, "2refine.bin", 0, 8), ("2d_2d2ref_d1_n8", 2, "2d2ref.bin", 1, 8),
               ("3d_multi_refine_n4", 3, "multi_refine.bin", 0, 4)]


def cycle_variant_cases():
    for name, D, mesh, div, n in CYCLE_CASES:
        base = np.load(os.path.join(HERE, name + ".npz"))
        f = base["rhs_f"]
        out = {"D": D, "n": n, "divide": div, "mesh": mesh}
        with tempfile.TemporaryDirectory() as tmp:
            t = lambda k: os.path.join(tmp, k)  # noqa: E731
            f.tofile(t("f"))
            for key, opt in CYCLE_VARIANTS.items():
                ref(D, mesh, div, n, "vcycle:%s:%s" % (t("f"), t("v")), solver="dft@" + opt)
                out["cycle_" + key] = np.fromfile(t("v"))
            # one block-Jacobi sweep with the shifted patch solver on every level, both solver implementations
            rng = np.random.default_rng(777)
            for l in range(int(base["nlevels"])):
                cells = base["L%d_in_u" % l].size
                u, ff = rng.standard_normal(cells), rng.standard_normal(cells)
                u.tofile(t("u"))
                ff.tofile(t("ff"))
                ref(D, mesh, div, n, "smooth:%d:%s:%s:%s" % (l, t("ff"), t("u"), t("s1")), solver="dft@lambda=-3.5")
                ref(D, mesh, div, n, "smooth:%d:%s:%s:%s" % (l, t("ff"), t("u"), t("s2")), solver="fftw@lambda=-3.5")
                out["L%d_in_u" % l], out["L%d_in_f" % l] = u, ff
                out["L%d_smooth_lambda" % l] = np.fromfile(t("s1"))
                out["L%d_smooth_lambda_fftw" % l] = np.fromfile(t("s2"))
        out["lambda"] = -3.5
        np.savez_compressed(os.path.join(HERE, name + "_cycles.npz"), **out)
        print(name + "_cycles", sorted(k for k in out if k.startswith("cycle_")))


def assembled_operator_case():
    """the reference's independent ASSEMBLED form of the operator (MatrixHelper::formCRSMatrix, per-side stencils of
    StencilHelper.h:298-538 incl. the coarse/fine weights; ref_gmg `matapply`) applied to the per-level inputs of the 3D
    Dirichlet and Neumann goldens: pins the matrix-free operator (and with it the ghost-fill weights) by a second
    implementation.  3D only: the 2D assembled stencil is a different discretisation."""
    out = {}
    for name, D, mesh, div, n in CASES + NEUMANN_CASES:
        if D != 3:
            continue
        base = np.load(os.path.join(HERE, name + ".npz"))
        solver = "dft-neumann" if name.endswith("_neumann") else "dft"
        with tempfile.TemporaryDirectory() as tmp:
            t = lambda k: os.path.join(tmp, k)  # noqa: E731
            for l in range(int(base["nlevels"])):
                base["L%d_in_u" % l].tofile(t("u"))
                ref(D, mesh, div, n, "matapply:%d:%s:%s" % (l, t("u"), t("m")), solver=solver)
                out["%s_L%d_matapply" % (name, l)] = np.fromfile(t("m"))
    np.savez_compressed(os.path.join(HERE, "3d_assembled_operator.npz"), **out)
    print("3d_assembled_operator", len(out), "vectors")


def main():
    if "--assembled-only" in sys.argv:
        return assembled_operator_case()
    if "--cycles-only" in sys.argv:
        return cycle_variant_cases()
    if "--neumann-init-only" in sys.argv:
        return neumann_init_case()
    if "--neumann-only" in sys.argv:
        return neumann_cases()
    neumann_init_case()
    neumann_cases()
    _base_cases()
    cycle_variant_cases()
    assembled_operator_case()


def _base_cases():
    for name, D, mesh, div, n in CASES:
        with tempfile.TemporaryDirectory() as tmp:
            t = lambda f: os.path.join(tmp, f)  # noqa: E731
            ref(D, mesh, div, n, "meta:" + t("meta"), "rhs:%s:%s" % (t("f"), t("e")))
            levels = go.read_ref_meta(t("meta"))
            out = {"D": D, "n": n, "divide": div, "mesh": mesh, "nlevels": len(levels)}
            for l, L in enumerate(levels):
                for k in ("ids", "refine_level", "parent_id", "orth_on_parent", "parent_idx", "neumann",
                          "starts", "spacings", "nbr_type", "nbr_ids", "nbr_idx", "orth_on_coarse"):
                    out["L%d_%s" % (l, k)] = getattr(L, k)
            out["rhs_f"] = np.fromfile(t("f"))
            out["rhs_exact"] = np.fromfile(t("e"))
            rng = np.random.default_rng(1234)
            for l, L in enumerate(levels):
                u = rng.standard_normal(L.cells)
                f = rng.standard_normal(L.cells)
                u.tofile(t("u"))
                f.tofile(t("ff"))
                cmds = ["apply:%d:%s:%s" % (l, t("u"), t("au")), "smooth:%d:%s:%s:%s" % (l, t("ff"), t("u"), t("su"))]
                out["L%d_in_u" % l] = u
                out["L%d_in_f" % l] = f
                if l + 1 < len(levels):
                    uc = rng.standard_normal(levels[l + 1].cells)
                    uc.tofile(t("uc"))
                    out["L%d_in_uc" % l] = uc
                    cmds += ["restrict:%d:%s:%s" % (l, t("u"), t("rc")),
                             "interp:%d:%s:%s:%s" % (l, t("uc"), t("u"), t("pi"))]
                ref(D, mesh, div, n, *cmds)
                out["L%d_apply" % l] = np.fromfile(t("au"))
                out["L%d_smooth" % l] = np.fromfile(t("su"))
                if l + 1 < len(levels):
                    out["L%d_restrict" % l] = np.fromfile(t("rc"))
                    out["L%d_interp" % l] = np.fromfile(t("pi"))
                # the FFTW-planned solver path (through the naive r2r stand-in) must agree too
                ref(D, mesh, div, n, "smooth:%d:%s:%s:%s" % (l, t("ff"), t("u"), t("su2")), solver="fftw")
                out["L%d_smooth_fftw" % l] = np.fromfile(t("su2"))
            ref(D, mesh, div, n, "vcycle:%s:%s" % (t("f"), t("v")),
                "vhist:%s:6:%s:%s" % (t("f"), t("vu"), t("vh")),
                "bicgstab:%s:1e-12:100:%s:%s" % (t("f"), t("bu"), t("bi")))
            out["vcycle"] = np.fromfile(t("v"))
            out["vhist_u"] = np.fromfile(t("vu"))
            out["vhist"] = np.fromfile(t("vh"))
            out["bicgstab_u"] = np.fromfile(t("bu"))
            out["bicgstab_info"] = np.fromfile(t("bi"))[:2]  # iterations, final relative residual
            np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
            print(name, [L.P for L in levels], "its", out["bicgstab_info"][0])


if __name__ == "__main__":
    main()

"""The real drop-in build (oracle/_ref/dropin_gmg = include/tgpu_thunderegg.hpp compiled against the reference's own
headers, oracle/dropin_driver.cpp): the reference's unmodified Init::initDirichlet, GMG::VCycle / GMG::WCycle and
BiCGStab<D>::solve drive the B200 kernels through adaptors that derive from the reference's Vector / Operator /
Smoother / Restrictor / Interpolator classes.  Results must equal the golden vectors the all-CPU reference produced."""
import json
import os
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import MESHES, ROOT, load_golden, rel_l2

BIN = os.path.join(ROOT, "oracle", "_ref", "dropin_gmg")


def test_dropin_binary_is_built_and_fails_loudly_without_gpu():
    if not os.path.exists(BIN):
        if not os.path.isdir("/root/reference/src/Thunderegg"):
            pytest.skip("needs the reference tree to build (prebuilt binary travels to the GPU box)")
        import build_native
        build_native.build()
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "dropin"])
    assert os.path.exists(BIN)
    import torch
    if not torch.cuda.is_available():
        with tempfile.TemporaryDirectory() as tmp:
            res = subprocess.run([BIN, "3", os.path.join(MESHES, "2refine.bin"), "0", "8", "-", tmp], capture_output=True, text=True)
            assert res.returncode != 0 and "no CPU fallback" in res.stderr


def _run(g, opts):
    with tempfile.TemporaryDirectory() as tmp:
        res = subprocess.run([BIN, str(int(g["D"])), os.path.join(MESHES, str(g["mesh"])), str(int(g["divide"])), str(int(g["n"])), opts, tmp],
                             capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stdout + res.stderr
        info = json.loads(res.stdout.strip().splitlines()[-1])
        return info, {k: np.fromfile(os.path.join(tmp, k + ".bin")) for k in
                      ("rhs_f", "rhs_exact", "vcycle_plugin", "vcycle_fused", "bicgstab_plugin", "bicgstab_fused")}


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["3d_2refine_n8", "2d_2d2ref_d1_n8", "3d_multi_refine_n4", "3d_2uni_n8"])
def test_reference_drivers_over_the_adaptors(name):
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/dropin_gmg not built")
    g = load_golden(name)
    info, out = _run(g, "-")
    # the reference's Init, copied host -> device through getLocalData / LocalDataManager and read back the same way
    assert np.array_equal(out["rhs_f"], g["rhs_f"]) and np.array_equal(out["rhs_exact"], g["rhs_exact"])
    assert abs(info["integral"] - info["integral_host"]) <= 1e-13 * max(1.0, abs(info["integral_host"]))
    # the reference's GMG::VCycle object over the adaptors, and the one-call fused cycle
    assert rel_l2(out["vcycle_plugin"], g["vcycle"]) < 1e-12
    assert rel_l2(out["vcycle_fused"], g["vcycle"]) < 1e-12
    # the reference's BiCGStab<D>::solve with either preconditioner: same iteration count and solution as the all-CPU run
    assert info["its_plugin"] == int(g["bicgstab_info"][0]) and info["its_fused"] == int(g["bicgstab_info"][0])
    assert rel_l2(out["bicgstab_plugin"], g["bicgstab_u"]) < 1e-10
    assert rel_l2(out["bicgstab_fused"], g["bicgstab_u"]) < 1e-10
    assert info["res_plugin"] < 1e-11 and info["res_fused"] < 1e-11
    assert info["foreign_vector_throw"] == 3  # reference convention for a foreign vector type


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["3d_2refine_n8", "2d_2d2ref_d1_n8"])
@pytest.mark.parametrize("key,opts", [("W", "W"), ("W_p2m2c2", "W,pre=2,mid=2,post=1,coarse=2"), ("V_max_levels2", "max_levels=2"),
                                      ("V_ppp2", "ppp=2")])
def test_reference_wcycle_and_level_options_over_the_adaptors(name, key, opts):
    """GMG::WCycle and the factory's max_levels / patches_per_proc rules, reference objects over the adaptors and the fused
    one-call form, against goldens of the all-CPU reference"""
    if not os.path.exists(BIN):
        pytest.skip("oracle/_ref/dropin_gmg not built")
    g, c = load_golden(name), load_golden(name + "_cycles")
    info, out = _run(g, opts)
    assert rel_l2(out["vcycle_plugin"], c["cycle_" + key]) < 1e-12
    assert rel_l2(out["vcycle_fused"], c["cycle_" + key]) < 1e-12

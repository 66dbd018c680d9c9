"""The C++ host side (include/tgpu_plugin.hpp + apps/steady.cpp) mirrors the reference's plugin
surface; these tests build it (CPU) and, on the GPU box, run the reference app's GMG path through
it: BiCGStab + cycle, plugin-granular (virtual Level/Smoother/... calls) and fused."""
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from conftest import MESHES, ROOT, load_golden, rel_l2

APP = os.path.join(ROOT, "apps", "steady")


def _build():
    import build_native
    build_native.build()
    return build_native.build_app()


def test_cpp_app_builds_and_fails_loudly_without_gpu():
    app = _build()
    assert os.path.exists(app)
    import torch
    if not torch.cuda.is_available():
        res = subprocess.run([app, "3", os.path.join(MESHES, "2refine.bin"), "0", "8"], capture_output=True, text=True)
        assert res.returncode != 0 and "no CPU fallback" in res.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [[], ["--plugin"], ["--cycle", "W"], ["--cycle", "W", "--plugin"], ["--max-levels", "2"]])
@pytest.mark.parametrize("name", ["3d_2refine_n8", "2d_2d2ref_d1_n8"])
def test_cpp_steady_matches_reference_solve(name, flags):
    g = load_golden(name)
    app = _build()
    with tempfile.TemporaryDirectory() as tmp:
        out = os.path.join(tmp, "u.bin")
        res = subprocess.run([app, str(int(g["D"])), os.path.join(MESHES, str(g["mesh"])), str(int(g["divide"])), str(int(g["n"])),
                              "--out", out] + flags, capture_output=True, text=True, check=True)
        its = int(re.search(r"Iterations: (\d+)", res.stdout).group(1))
        assert float(re.search(r"Residual: (\S+)", res.stdout).group(1)) < 1e-11
        if not [f for f in flags if f != "--plugin"]:  # the reference's default cycle: same Krylov trajectory as the golden solve
            assert its == int(g["bicgstab_info"][0])
            assert rel_l2(np.fromfile(out), g["bicgstab_u"]) < 1e-10
        else:  # other preconditioners converge to the same discrete solution
            assert rel_l2(np.fromfile(out), g["bicgstab_u"]) < 1e-9

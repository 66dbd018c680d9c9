"""CPU-side checks of the product library: it loads, exports every symbol include/tgpu.h declares,
fails loudly without a GPU, and its host mesh ingest reproduces the reference's per-level
metadata (tests/golden/*.npz, produced by the reference itself) bit for bit."""
import os
import re

import numpy as np
import pytest

import pressurepoissonsolver_b200 as pps
from conftest import GOLDEN_CASES, MESHES, ROOT, load_golden


def test_header_symbols_all_exported():
    hdr = open(os.path.join(ROOT, "include", "tgpu.h")).read()
    declared = set(re.findall(r"\b(tgpu_[a-z0-9_]+)\s*\(", hdr))
    assert declared == set(pps.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(pps.lib, name), name


def test_version_string():
    assert b"sm_100a" in pps.lib.tgpu_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(pps.TgpuError, match="no CPU fallback"):
        pps.Context(0)


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_mesh_levels_match_reference(name):
    g = load_golden(name)
    D, n = int(g["D"]), int(g["n"])
    mesh = pps.Mesh.load(os.path.join(MESHES, str(g["mesh"])), D)
    mesh.refine_leaves(int(g["divide"]))
    levels = mesh.level_arrays(n)
    assert len(levels) == int(g["nlevels"])
    for l, L in enumerate(levels):
        for k in ("ids", "refine_level", "parent_id", "orth_on_parent", "nbr_type", "nbr_idx", "orth_on_coarse",
                  "starts", "spacings"):
            assert np.array_equal(L[k], g["L%d_%s" % (l, k)]), (l, k)
        if l + 1 < len(levels):
            assert np.array_equal(L["parent_idx"], g["L%d_parent_idx" % l])
    mesh.close()


def test_uniform_generator_equals_shipped_uniform_trees():
    """tgpu_mesh_uniform(3, L) must give the same level tables as the reference's Luni.bin files."""
    for L, fname in ((2, "2uni.bin"), (3, "3uni.bin"), (4, "4uni.bin")):
        a = pps.Mesh.uniform(3, L).level_arrays(4)
        b = pps.Mesh.load(os.path.join(MESHES, fname), 3).level_arrays(4)
        assert len(a) == len(b) == L
        for la, lb in zip(a, b):
            for k in ("nbr_type", "nbr_idx", "orth_on_coarse", "orth_on_parent", "parent_idx", "starts", "spacings"):
                assert np.array_equal(la[k], lb[k]), (fname, k)


def test_refine_counts_like_reference_octtree_test():
    """test/OctTree.cpp:5-171: refining the single-node tree once gives 9 nodes, twice 73."""
    m = pps.Mesh.load(os.path.join(MESHES, "1uni.bin"), 3)
    assert m.info()[2] == 1
    m.refine_leaves()
    assert m.info()[2] == 9
    m.refine_leaves()
    assert m.info()[2] == 73 and m.info()[1] == 3


def test_bad_arguments_report_errors():
    with pytest.raises(pps.TgpuError):
        pps.Mesh.load("/nonexistent/mesh.bin", 3)
    with pytest.raises(pps.TgpuError):
        pps.Mesh.uniform(3, 2).extract_levels(5)  # odd n
def test_corrupt_mesh_files_are_rejected(tmp_path):
    """a truncated file, a 3D file opened as 2D and dangling ids are I/O errors, not out-of-bounds accesses"""
    src = open(os.path.join(MESHES, "2refine.bin"), "rb").read()
    bad = tmp_path / "trunc.bin"
    bad.write_bytes(src[:len(src) - 40])
    with pytest.raises(pps.TgpuError):
        pps.Mesh.load(str(bad), 3)
    with pytest.raises(pps.TgpuError):
        pps.Mesh.load(os.path.join(MESHES, "2refine.bin"), 2)
    b = bytearray(src)
    b[8 + 8:8 + 12] = (123456).to_bytes(4, "little")  # the root's parent id -> dangling
    bad2 = tmp_path / "dangling.bin"
    bad2.write_bytes(bytes(b))
    with pytest.raises(pps.TgpuError):
        pps.Mesh.load(str(bad2), 3)



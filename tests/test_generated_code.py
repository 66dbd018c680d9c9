"""The FFT-based DST-IV header (csrc/dst4_fast.cuh) is generated: the committed file must be what the generator
emits, and the emitted arithmetic must reproduce the dense DST-IV matrix (the dense half of the reference's
DST-II / DST-III patch transforms, DftPatchSolver.h:262-281)."""
import importlib.util
import math
import os
import random

from conftest import ROOT


def _gen():
    spec = importlib.util.spec_from_file_location("gen_dst4", os.path.join(ROOT, "tools", "gen_dst4.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_generated_dst4_matches_dense_matrix():
    gen = _gen()
    for N in (8, 16, 32):
        body = gen.gen(N)
        assert gen.check(N, body) < 1e-14 * N / 8
        # an independent evaluation on another input
        rnd = random.Random(N)
        x = [rnd.uniform(-2, 2) for _ in range(N)]
        env = {"x": x, "y": [0.0] * N, "fma": lambda a, b, c: a * b + c}
        for ln in body:
            exec(ln.replace("const double ", "").rstrip(";"), env)
        for k in range(N):
            ref = sum(x[n] * math.sin(math.pi * (2 * k + 1) * (2 * n + 1) / (4 * N)) for n in range(N))
            assert abs(env["y"][k] - ref) < 1e-13 * N / 8


def test_committed_header_is_generator_output():
    gen = _gen()
    text = open(os.path.join(ROOT, "pressurepoissonsolver_b200", "csrc", "dst4_fast.cuh")).read()
    for N in (8, 16, 32):
        for ln in gen.gen(N):
            assert ln in text, (N, ln)

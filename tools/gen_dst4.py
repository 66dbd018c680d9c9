"""Generates pressurepoissonsolver_b200/csrc/dst4_fast.cuh: straight-line DST-IV kernels of size 8, 16 and 32,
    Y_k = sum_n x_n sin(pi (2k+1)(2n+1) / (4N)),
the dense half of the symmetric split of the DST-II / DST-III patch transforms (kernels.cuh, Dst2 / Dst3;
DftPatchSolver.h:262-281 in the reference).  Algorithm: DST-IV(x)_k = (-1)^k DCT-IV(reversed x)_k and the DCT-IV
through one complex FFT of size N/2 with pre- and post-twiddles:
    z_m = x'_{2m} + i x'_{N-1-2m},  t_m = z_m e^{-i pi (4m+1)/(4N)},  T = FFT_{N/2}(t),  w_k = e^{-i pi k/N} T_k,
    C_{2k} = Re w_k,  C_{N-1-2k} = -Im w_k.
The script records the arithmetic symbolically, emits CUDA, and checks the emitted code against the dense matrix.
Usage: python tools/gen_dst4.py  (rewrites the header; the header is committed)."""
import math
import os
import random

lines = []
counter = [0]


def new():
    counter[0] += 1
    return "t%d" % counter[0]


def lit(c):
    return repr(float(c))


class V:
    """a real value held in a named temporary (or an input)"""

    def __init__(self, name):
        self.n = name

    def __add__(self, o):
        r = new()
        lines.append("const double %s = %s + %s;" % (r, self.n, o.n))
        return V(r)

    def __sub__(self, o):
        r = new()
        lines.append("const double %s = %s - %s;" % (r, self.n, o.n))
        return V(r)

    def scale(self, c):
        r = new()
        lines.append("const double %s = %s * %s;" % (r, lit(c), self.n))
        return V(r)

    def neg(self):
        r = new()
        lines.append("const double %s = -%s;" % (r, self.n))
        return V(r)


def fma(c, a, b):
    r = new()
    lines.append("const double %s = fma(%s, %s, %s);" % (r, lit(c), a.n, b.n))
    return V(r)


def cmul(re, im, c, s):
    """(re + i im) (c + i s) with constants c, s: 2 multiplications + 2 fused multiply-adds"""
    if abs(s) < 1e-300 and abs(c - 1) < 1e-300:
        return re, im
    p = im.scale(-s)
    q = re.scale(s)
    return fma(c, re, p), fma(c, im, q)


def fft(re, im):
    """decimation-in-frequency radix-2 complex FFT, kernel e^{-2 pi i nk/n}; natural-order output"""
    n = len(re)
    if n == 1:
        return re, im
    h = n // 2
    ar, ai, br, bi = [], [], [], []
    for k in range(h):
        ar.append(re[k] + re[k + h])
        ai.append(im[k] + im[k + h])
        dr, di = re[k] - re[k + h], im[k] - im[k + h]
        ang = -2 * math.pi * k / n
        c, s = math.cos(ang), math.sin(ang)
        if k == 0:
            pass
        elif 4 * k == n:  # multiply by -i: (dr + i di)(-i) = di - i dr
            dr, di = di, dr.neg()
        elif 8 * k == n or 8 * k == 3 * n:  # |c| == |s|: two additions + two multiplications
            m = abs(c)
            if 8 * k == n:  # (1 - i)/sqrt2: re = (dr + di) m, im = (di - dr) m
                dr, di = (dr + di).scale(m), (di - dr).scale(m)
            else:  # (-1 - i)/sqrt2: re = (di - dr) m, im = -(dr + di) m
                dr, di = (di - dr).scale(m), (dr + di).scale(-m)
        else:
            dr, di = cmul(dr, di, c, s)
        br.append(dr)
        bi.append(di)
    er, ei = fft(ar, ai)
    orr, oi = fft(br, bi)
    outr, outi = [None] * n, [None] * n
    for k in range(h):
        outr[2 * k], outi[2 * k] = er[k], ei[k]
        outr[2 * k + 1], outi[2 * k + 1] = orr[k], oi[k]
    return outr, outi


def gen(N):
    global lines
    lines = []
    counter[0] = 0
    x = [V("x[%d]" % n) for n in range(N)]
    xr = x[::-1]  # DST-IV(x)_k = (-1)^k DCT-IV(reversed x)_k
    H = N // 2
    tr, ti = [], []
    for m in range(H):
        ang = -math.pi * (4 * m + 1) / (4 * N)
        a, b = cmul(xr[2 * m], xr[N - 1 - 2 * m], math.cos(ang), math.sin(ang))
        tr.append(a)
        ti.append(b)
    Tr, Ti = fft(tr, ti)
    out = [None] * N
    for k in range(H):
        ang = -math.pi * k / N
        c, s = math.cos(ang), math.sin(ang)
        # C_{2k} = Re w, C_{N-1-2k} = -Im w; Y_j = (-1)^j C_j: fold the signs into the constants
        j0, j1 = 2 * k, N - 1 - 2 * k
        s0 = 1.0 if j0 % 2 == 0 else -1.0
        s1 = -1.0 if j1 % 2 == 0 else 1.0  # includes the minus of -Im
        if k == 0:
            out[j0] = Tr[0] if s0 > 0 else Tr[0].neg()
            out[j1] = Ti[0] if s1 > 0 else Ti[0].neg()
        else:
            # Re w = c Tr - s Ti, Im w = c Ti + s Tr
            out[j0] = fma(s0 * c, Tr[k], Ti[k].scale(-s0 * s))
            out[j1] = fma(s1 * c, Ti[k], Tr[k].scale(s1 * s))
    body = list(lines)
    for j in range(N):
        body.append("y[%d] = %s;" % (j, out[j].n))
    return body


def check(N, body):
    rnd = random.Random(5)
    x = [rnd.uniform(-1, 1) for _ in range(N)]
    y = [0.0] * N
    env = {"x": x, "y": y, "fma": lambda a, b, c: a * b + c}
    for ln in body:
        exec(ln.replace("const double ", "").rstrip(";"), env)
    ref = [sum(x[n] * math.sin(math.pi * (2 * k + 1) * (2 * n + 1) / (4 * N)) for n in range(N)) for k in range(N)]
    err = max(abs(a - b) for a, b in zip(y, ref))
    assert err < 1e-13, (N, err)  # sums of N terms of size <= 1: a few ulps of N
    return err


def main():
    out = ["// dst4_fast.cuh - GENERATED by tools/gen_dst4.py, do not edit.",
           "// Straight-line DST-IV of size 8, 16 and 32, Y_k = sum_n x_n sin(pi (2k+1)(2n+1) / (4N)): the dense half of the",
           "// symmetric split of the DST-II / DST-III patch transforms (Dst2 / Dst3 in kernels.cuh), computed through a",
           "// complex FFT of size N/2 with pre- and post-twiddles instead of an N x N product.",
           "#pragma once", "namespace tgpu", "{"]
    for N in (8, 16, 32):
        body = gen(N)
        err = check(N, body)
        nops = sum(1 for b in body if b.startswith("const double"))
        out.append("// %d fp64 instructions instead of %d (checked against the dense matrix: max error %.1e)" % (nops, N * N, err))
        out.append("__device__ __forceinline__ void dst4_%d(const double (&x)[%d], double (&y)[%d])" % (N, N, N))
        out.append("{")
        out += ["\t" + b for b in body]
        out.append("}")
        print("N=%d: %d instructions, max error %.2e" % (N, nops, err))
    out.append("} // namespace tgpu")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "pressurepoissonsolver_b200", "csrc", "dst4_fast.cuh")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""Generate straight-line, in-register complex-FP64 DFT codelets for the sm_100a FFT kernels.

Writes sternheimergw_b200/csrc/fft_codelets.h with
    template <int R> SGW_HD void dft_fwd(double* re, double* im);      // X[k] = sum_j x[j] e^{-2 pi i jk/R}
for R in RADICES, operating in place on per-thread arrays (fully unrolled -> registers).
The unscaled inverse is obtained for free by swapping the roles of re/im:  dft_fwd<R>(im, re).

Composite radices are built by Cooley-Tukey from the prime butterflies 2, 3, 5 (and 4) with the
inner twiddles folded in as literal constants (multiplications by 1, -1, +-i and (+-1+-i)/sqrt2 are
strength-reduced).  The header is plain C++ (SGW_HD = __host__ __device__ under nvcc) so that the
same code is unit-tested on the CPU (tests/test_fft_core_cpu.py).
"""
import math
from pathlib import Path

RADICES = [2, 3, 4, 5, 6, 8, 9, 10, 12, 15, 16, 18, 20, 24, 25, 27, 30, 32, 7, 11, 14, 21, 22, 28]
BIG = {18, 20, 24, 25, 27, 30, 32, 7, 11, 14, 21, 22, 28}      # dispatched by the run_*_big functions of fft_core.h
OUT = Path(__file__).resolve().parent.parent / "sternheimergw_b200" / "csrc" / "fft_codelets.h"


class Emitter:
    def __init__(self):
        self.lines = []
        self.n = 0

    def tmp(self, expr):
        self.n += 1
        name = f"t{self.n}"
        self.lines.append(f"  const double {name} = {expr};")
        return name

    # complex values are (re_name, im_name)
    def add(self, a, b):
        return (self.tmp(f"{a[0]} + {b[0]}"), self.tmp(f"{a[1]} + {b[1]}"))

    def sub(self, a, b):
        return (self.tmp(f"{a[0]} - {b[0]}"), self.tmp(f"{a[1]} - {b[1]}"))

    def scale(self, a, c):
        return (self.tmp(f"{c!r} * {a[0]}"), self.tmp(f"{c!r} * {a[1]}"))

    def mul_neg_i(self, a):      # a * (-i) = (im, -re)
        return (a[1], self.tmp(f"-{a[0]}"))

    def mul_pos_i(self, a):      # a * (+i) = (-im, re)
        return (self.tmp(f"-{a[1]}"), a[0])

    def fma_c(self, a, c, b):    # a + c*b  (c real constant)
        return (self.tmp(f"fma({c!r}, {b[0]}, {a[0]})"), self.tmp(f"fma({c!r}, {b[1]}, {a[1]})"))

    def mul_w(self, a, num, den):
        """a * exp(-2 pi i num/den)"""
        num %= den
        if num == 0:
            return a
        if 4 * num == den:
            return self.mul_neg_i(a)
        if 2 * num == den:
            return (self.tmp(f"-{a[0]}"), self.tmp(f"-{a[1]}"))
        if 4 * num == 3 * den:
            return self.mul_pos_i(a)
        ang = -2.0 * math.pi * num / den
        c, s = math.cos(ang), math.sin(ang)
        if (8 * num) % den == 0:          # odd multiples of pi/4: |c| == |s|
            h = math.sqrt(0.5)
            sc, ss = (1 if c > 0 else -1), (1 if s > 0 else -1)
            # (xr + i xi)(c + i s) = (xr c - xi s) + i (xr s + xi c)
            re = self.tmp(f"{h!r} * ({'' if sc > 0 else '-'}{a[0]} {'-' if ss > 0 else '+'} {a[1]})")
            im = self.tmp(f"{h!r} * ({'' if ss > 0 else '-'}{a[0]} {'+' if sc > 0 else '-'} {a[1]})")
            return (re, im)
        re = self.tmp(f"fma({c!r}, {a[0]}, {(-s)!r} * {a[1]})")
        im = self.tmp(f"fma({s!r}, {a[0]}, {c!r} * {a[1]})")
        return (re, im)


def dft(e: Emitter, x):
    """forward DFT of the list of complex values x; returns list in natural order"""
    R = len(x)
    if R == 1:
        return x
    if R == 2:
        return [e.add(x[0], x[1]), e.sub(x[0], x[1])]
    if R == 3:
        t1 = e.add(x[1], x[2])
        d = e.sub(x[1], x[2])
        X0 = e.add(x[0], t1)
        m1 = e.fma_c(x[0], -0.5, t1)
        s = math.sqrt(3.0) / 2.0
        # -i * s * d = (s*d_im, -s*d_re)
        r = e.tmp(f"{s!r} * {d[1]}")
        i = e.tmp(f"{s!r} * {d[0]}")
        X1 = (e.tmp(f"{m1[0]} + {r}"), e.tmp(f"{m1[1]} - {i}"))
        X2 = (e.tmp(f"{m1[0]} - {r}"), e.tmp(f"{m1[1]} + {i}"))
        return [X0, X1, X2]
    if R == 4:
        a, b, c, d = x
        s02, d02 = e.add(a, c), e.sub(a, c)
        s13, d13 = e.add(b, d), e.sub(b, d)
        X0, X2 = e.add(s02, s13), e.sub(s02, s13)
        # X1 = d02 - i d13 ; X3 = d02 + i d13
        X1 = (e.tmp(f"{d02[0]} + {d13[1]}"), e.tmp(f"{d02[1]} - {d13[0]}"))
        X3 = (e.tmp(f"{d02[0]} - {d13[1]}"), e.tmp(f"{d02[1]} + {d13[0]}"))
        return [X0, X1, X2, X3]
    if R == 5:
        a, b, c, d, f = x
        t1, t2 = e.add(b, f), e.add(c, d)
        t3, t4 = e.sub(b, f), e.sub(c, d)
        c1, c2 = math.cos(2 * math.pi / 5), math.cos(4 * math.pi / 5)
        s1, s2 = math.sin(2 * math.pi / 5), math.sin(4 * math.pi / 5)
        X0 = e.add(a, e.add(t1, t2))
        A1 = e.fma_c(e.fma_c(a, c1, t1), c2, t2)
        A2 = e.fma_c(e.fma_c(a, c2, t1), c1, t2)
        B1 = (e.tmp(f"fma({s1!r}, {t3[0]}, {s2!r} * {t4[0]})"), e.tmp(f"fma({s1!r}, {t3[1]}, {s2!r} * {t4[1]})"))
        B2 = (e.tmp(f"fma({s2!r}, {t3[0]}, {(-s1)!r} * {t4[0]})"), e.tmp(f"fma({s2!r}, {t3[1]}, {(-s1)!r} * {t4[1]})"))
        # X1 = A1 - i B1 ; X4 = A1 + i B1 ; X2 = A2 - i B2 ; X3 = A2 + i B2      (-iB = (B_im, -B_re))
        X1 = (e.tmp(f"{A1[0]} + {B1[1]}"), e.tmp(f"{A1[1]} - {B1[0]}"))
        X4 = (e.tmp(f"{A1[0]} - {B1[1]}"), e.tmp(f"{A1[1]} + {B1[0]}"))
        X2 = (e.tmp(f"{A2[0]} + {B2[1]}"), e.tmp(f"{A2[1]} - {B2[0]}"))
        X3 = (e.tmp(f"{A2[0]} - {B2[1]}"), e.tmp(f"{A2[1]} + {B2[0]}"))
        return [X0, X1, X2, X3, X4]
    if R in (7, 11, 13):
        # odd prime p: X_k = x_0 + sum_{j=1..h} [ (x_j + x_{p-j}) cos(2 pi j k / p) - i (x_j - x_{p-j}) sin(2 pi j k / p) ], h = (p-1)/2;
        # X_{p-k} is the same with the sign of the sine part flipped
        p, h = R, (R - 1) // 2
        sp = [e.add(x[j], x[p - j]) for j in range(1, h + 1)]
        dm = [e.sub(x[j], x[p - j]) for j in range(1, h + 1)]
        tot = x[0]
        for t in sp:
            tot = e.add(tot, t)
        X = [None] * p
        X[0] = tot
        for k in range(1, h + 1):
            A = x[0]
            B = None
            for j in range(1, h + 1):
                c = math.cos(2 * math.pi * j * k / p)
                sn = math.sin(2 * math.pi * j * k / p)
                A = e.fma_c(A, c, sp[j - 1])
                B = e.scale(dm[j - 1], sn) if B is None else e.fma_c(B, sn, dm[j - 1])
            # X_k = A - i B ; X_{p-k} = A + i B        (-iB = (B_im, -B_re))
            X[k] = (e.tmp(f"{A[0]} + {B[1]}"), e.tmp(f"{A[1]} - {B[0]}"))
            X[p - k] = (e.tmp(f"{A[0]} - {B[1]}"), e.tmp(f"{A[1]} + {B[0]}"))
        return X
    # composite: R = a*b, j = j2 + b*j1, k = k1 + a*k2
    for a in (4, 2, 3, 5, 7, 11):
        if R % a == 0 and R // a > 1:
            break
    else:
        raise ValueError(R)
    b = R // a
    Y = [[None] * a for _ in range(b)]
    for j2 in range(b):
        sub = dft(e, [x[j2 + b * j1] for j1 in range(a)])
        for k1 in range(a):
            Y[j2][k1] = e.mul_w(sub[k1], j2 * k1, R)
    X = [None] * R
    for k1 in range(a):
        col = dft(e, [Y[j2][k1] for j2 in range(b)])
        for k2 in range(b):
            X[k1 + a * k2] = col[k2]
    return X


def main():
    out = ["// GENERATED by tools/gen_codelets.py -- do not edit.",
           "// In-register complex-FP64 DFT codelets (forward, unscaled).  Inverse: dft_fwd<R>(im, re).",
           "#pragma once", "#include <math.h>", "#ifndef SGW_HD", "#ifdef __CUDACC__",
           "#define SGW_HD __host__ __device__ __forceinline__", "#else", "#define SGW_HD inline", "#endif", "#endif",
           "", "namespace sgw {", "", "template <int R> SGW_HD void dft_fwd(double* re, double* im);", "",
           "template <> SGW_HD void dft_fwd<1>(double*, double*) {}", ""]
    for R in RADICES:
        e = Emitter()
        x = [(f"re[{j}]", f"im[{j}]") for j in range(R)]
        # load inputs into named temporaries first (outputs overwrite the arrays)
        xin = [(e.tmp(a), e.tmp(b)) for a, b in x]
        X = dft(e, xin)
        out.append(f"template <> SGW_HD void dft_fwd<{R}>(double* re, double* im) {{")
        out.extend(e.lines)
        for k in range(R):
            out.append(f"  re[{k}] = {X[k][0]}; im[{k}] = {X[k][1]};")
        out.append("}")
        out.append("")
    # the radices above 16 are dispatched by separate functions (fft_core.h run_*_big) so that their register pressure does
    # not touch the code generated for the common ones
    out.append("#define SGW_FOR_EACH_RADIX(X) " + " ".join(f"X({r})" for r in [1] + [r for r in RADICES if r not in BIG]))
    out.append("#define SGW_FOR_EACH_BIG_RADIX(X) " + " ".join(f"X({r})" for r in RADICES if r in BIG))
    out.append("")
    out.append("}  // namespace sgw")
    import sys
    dst = Path(sys.argv[1]) if len(sys.argv) > 1 else OUT        # optional output path (tests regenerate into a scratch file)
    dst.parent.mkdir(parents=True, exist_ok=True)
    dst.write_text("\n".join(out) + "\n")
    print("wrote", dst, len(out), "lines")


if __name__ == "__main__":
    main()

"""The device line-FFT stages (sternheimergw_b200/csrc/fft_core.h, shared verbatim by the sm_100a kernels) compiled
for the host and checked against numpy.fft: every length the BASELINE grids use, both layouts, several simulated
thread counts (the task -> thread mapping must not matter), and the fused inverse -> x v(r) -> forward middle stage
of the local-potential product ([QE] vloc_psi_k semantics: invfft unscaled, fwfft scaled by the caller)."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
SRC = ROOT / "tests" / "cpu_harness" / "fft_core_test.cpp"


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    exe = tmp_path_factory.mktemp("fftcore") / "fft_core_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-o", str(exe), str(SRC)], check=True)
    return exe


def _run(exe, n, nlines, direction, nthreads, layout, x, v=None):
    txt = f"{n} {nlines} {direction} {nthreads} {layout}\n"
    txt += "\n".join(f"{float(z.real)!r} {float(z.imag)!r}" for z in x.ravel()) + "\n"
    if v is not None:
        txt += "\n".join(repr(float(q)) for q in v.ravel()) + "\n"
    out = subprocess.run([str(exe)], input=txt, capture_output=True, text=True, check=True).stdout.split("\n")
    if out[0].startswith("NOPLAN"):
        return None, None
    r1, r2 = map(int, out[0].split())
    vals = np.array([[float(a) for a in ln.split()] for ln in out[1:] if ln.strip()])
    return (vals[:, 0] + 1j * vals[:, 1]).reshape(nlines, n), (r1, r2)


def _perm(n, r1, r2):
    pos = np.arange(n)
    return pos // r2 + r1 * (pos % r2)          # position -> natural index (fft_core.h perm_index)


@pytest.mark.parametrize("n", [15, 18, 20, 24, 60, 72, 8, 9, 10, 36, 45, 48, 64, 96, 100, 120, 144, 225, 256,
                               # lengths that need the radices 18..32 (125, 162, 200, 216 ... were rejected)
                               125, 162, 200, 216, 243, 250, 270, 288, 320, 375, 384, 405, 480, 512, 625, 768, 900, 960,
                               # lengths with one factor 7 and/or 11 (QE's good_fft_order hands them out)
                               7, 11, 14, 21, 22, 28, 42, 44, 56, 63, 66, 70, 77, 84, 88, 105, 112, 154, 231, 308])
@pytest.mark.parametrize("layout", [0, 1])
def test_line_fft_matches_numpy(harness, n, layout):
    rng = np.random.default_rng(n + layout)
    nlines = 5
    x = rng.standard_normal((nlines, n)) + 1j * rng.standard_normal((nlines, n))
    for nthreads in (1, 7, 64):
        y, (r1, r2) = _run(harness, n, nlines, +1, nthreads, layout, x)
        ref = np.fft.ifft(x, axis=1) * n                       # invfft: unscaled, e^{+i}
        assert np.abs(y - ref[:, _perm(n, r1, r2)]).max() < 1e-12 * n
        xin = x[:, _perm(n, r1, r2)]                            # permuted-order input -> natural-order output
        z, _ = _run(harness, n, nlines, -1, nthreads, layout, xin)
        assert np.abs(z - np.fft.fft(x, axis=1)).max() < 1e-12 * n


def test_every_5_smooth_length_up_to_960_has_a_plan(harness):
    def smooth(n):
        for p in (2, 3, 5):
            while n % p == 0:
                n //= p
        return n == 1
    for n in range(1, 961):
        if smooth(n):
            y, plan = _run(harness, n, 1, +1, 1, 0, np.ones((1, n), complex))
            assert y is not None and plan[0] * plan[1] == n and max(plan) <= 32, n


def test_unsupported_length_has_no_plan(harness):
    for n in (13, 26, 221):
        y, _ = _run(harness, n, 1, +1, 1, 0, np.zeros((1, n), complex))
        assert y is None


@pytest.mark.parametrize("n", [15, 18, 20, 24, 60, 72, 12, 16, 125, 200, 216, 512, 14, 28, 42, 77])
def test_fused_vloc_middle_stage(harness, n):
    rng = np.random.default_rng(100 + n)
    nlines = 6
    x = rng.standard_normal((nlines, n)) + 1j * rng.standard_normal((nlines, n))
    v_nat = rng.standard_normal((nlines, n))
    r = np.fft.ifft(x, axis=1) * n
    ref = np.fft.fft(v_nat * r, axis=1)
    for nthreads in (1, 5, 32):
        _, (r1, r2) = _run(harness, n, 1, +1, 1, 0, x[:1])
        v_perm = v_nat[:, _perm(n, r1, r2)]                     # the library stores v(r) pre-permuted
        y, _ = _run(harness, n, nlines, 0, nthreads, 0, x, v_perm)
        assert np.abs(y - ref).max() < 1e-12 * n * n


def test_tracked_codelets_are_what_the_generator_writes(tmp_path):
    """csrc/fft_codelets.h is generated (tools/gen_codelets.py); the tracked copy must be the generator's output."""
    import sys
    out = tmp_path / "fft_codelets.h"
    subprocess.run([sys.executable, str(ROOT / "tools" / "gen_codelets.py"), str(out)], check=True, capture_output=True)
    assert out.read_text() == (ROOT / "sternheimergw_b200" / "csrc" / "fft_codelets.h").read_text()

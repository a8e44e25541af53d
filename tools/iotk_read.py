#!/usr/bin/env python3
"""Convert the reference's solver golden vector into a small .npz fixture.

Source (read-only, only available in the build container):
  /root/reference/algo/linear_solver/test/lin_prob.xml.bz2
written by phys/green/src/green.f90:438-446 (green_solver_debug) and read by
algo/linear_solver/test/linear_solver.pf:44-97 (linear_problem_read).

The file is an iotk *binary* XML: a sequence of Fortran unformatted records.
Each <TAG type=.. size=.. kind=..> tag record is followed by a data record
  [int32 reclen][int32 iotk header][payload][int32 reclen]
so the payload starts 12 bytes after the end of the tag line.

Run (in the build container only):
  python tools/iotk_read.py            # writes tests/golden/lin_prob.npz
The operator is real symmetric to 1.4e-16, so only Re(A) is stored (the
imaginary part's max-abs is recorded and asserted to be < 1e-15).
"""
import bz2
import re
import sys
from pathlib import Path

import numpy as np

SRC = Path("/root/reference/algo/linear_solver/test/lin_prob.xml.bz2")
DST = Path(__file__).resolve().parent.parent / "tests" / "golden" / "lin_prob.npz"

_DT = {"integer": "<i4", "complex": "<c16", "real": "<f8"}


def getdat(buf: bytes, tag: bytes) -> np.ndarray:
    m = re.search(rb"<" + tag + rb' type="(\w+)" size="(\d+)" kind="(\d+)">\n', buf)
    if m is None:
        raise KeyError(tag)
    typ, size, kind = m.group(1).decode(), int(m.group(2)), int(m.group(3))
    nbytes = size * kind * (2 if typ == "complex" else 1)
    off = m.end() + 12
    return np.frombuffer(buf[off:off + nbytes], dtype=_DT[typ]).copy()


def read_lin_prob(path: Path = SRC):
    buf = bz2.decompress(path.read_bytes())
    n = int(getdat(buf, b"DIMENSION")[0])
    ns = int(getdat(buf, b"NUMBER_SHIFT")[0])
    sigma = getdat(buf, b"LIST_SHIFT")
    A = getdat(buf, b"LINEAR_OPERATOR").reshape(n, n, order="F")
    b = getdat(buf, b"RIGHT_HAND_SIDE")
    x_bad = getdat(buf, b"INCORRECT_SOLUTION").reshape(n, ns, order="F")
    assert sigma.shape == (ns,) and b.shape == (n,)
    return n, ns, sigma, A, b, x_bad


def main():
    n, ns, sigma, A, b, x_bad = read_lin_prob()
    imag_max = float(np.abs(A.imag).max())
    assert imag_max < 1e-15, imag_max
    assert np.abs(A - A.conj().T).max() < 1e-15
    ev = np.linalg.eigvalsh(A.real)
    print(f"n={n} ns={ns} eig[0..4]={ev[:4]} eig[-1]={ev[-1]} |Im A|max={imag_max:.2e}")
    print("b nonzeros:", np.flatnonzero(b), b[np.flatnonzero(b)])
    DST.parent.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(DST, n=n, ns=ns, sigma=sigma, A_real=A.real, A_imag_max=imag_max,
                        b=b, x_bad=x_bad.astype(np.complex64))
    print("wrote", DST, DST.stat().st_size, "bytes")


if __name__ == "__main__":
    sys.exit(main())

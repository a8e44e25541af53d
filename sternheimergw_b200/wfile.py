"""On-disk W(q) records of the reference, so that the GPU path can exchange the screened
Coulomb interaction with an unmodified ``gw.x``.

Layout (``algo/io/src/opengwfil.f90:51-55``, written by ``davcio(scrcoul_g, lrcoul, iuncoul, iq, +1)`` at
``phys/coul/src/do_stern.f90:236``, read back at ``phys/corr/src/sigma.f90:306-331`` with
``ACCESS='direct', RECL = byte_real * lrcoul``): a Fortran direct-access file ``<outdir>/_gw0/<prefix>.coul1`` whose record
``iq`` (1-based q index) holds ``scrcoul_g(num_g_corr, num_g_corr, nfs)`` as ``lrcoul = 2 * num_g_corr**2 * nfs`` reals of
8 bytes, i.e. the column-major COMPLEX(dp) array, without record markers (direct access), little endian.
After ``invert_epsilon`` (direct solver) the array is eps^-1 - 1 (``invert_epsilon.f90:84-88``).

Pure host I/O: no GPU involved, numpy only.
"""
from __future__ import annotations

import os

import numpy as np

BYTE_REAL = 8      # sigma.f90:312 byte_real


def lrcoul(num_g_corr: int, nfs: int) -> int:
    """Record length in reals (opengwfil.f90:53)."""
    return 2 * num_g_corr * num_g_corr * nfs


def coul_filename(outdir: str, prefix: str, filcoul: str = "coul") -> str:
    """<tmp_dir_coul><prefix>.<filcoul>1 with tmp_dir_coul = <outdir>/_gw0/ (sigma.f90:308)."""
    return os.path.join(outdir, "_gw0", f"{prefix}.{filcoul}1")


def write_w_record(path: str, iq: int, scrcoul_g: np.ndarray) -> None:
    """davcio(scrcoul_g, lrcoul, iuncoul, iq, +1): write record iq (1-based) of the direct-access file."""
    a = np.asarray(scrcoul_g, dtype=np.complex128)
    if a.ndim != 3 or a.shape[0] != a.shape[1]:
        raise ValueError("scrcoul_g must be (num_g_corr, num_g_corr, nfs)")
    if iq < 1:
        raise ValueError("record index iq is 1-based")
    rec = lrcoul(a.shape[0], a.shape[2]) * BYTE_REAL
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    mode = "r+b" if os.path.exists(path) else "w+b"
    with open(path, mode) as f:
        f.seek((iq - 1) * rec)
        f.write(np.asfortranarray(a).astype("<c16").tobytes(order="F"))


def read_w_record(path: str, iq: int, num_g_corr: int, nfs: int) -> np.ndarray:
    """davcio(coulomb, lrcoul, iuncoul, iq, -1): read record iq (1-based); raises if the record is not there."""
    rec = lrcoul(num_g_corr, nfs) * BYTE_REAL
    with open(path, "rb") as f:
        f.seek((iq - 1) * rec)
        buf = f.read(rec)
    if len(buf) != rec:
        raise IOError(f"record {iq} of {path} is incomplete ({len(buf)} of {rec} bytes)")       # davcio: error reading
    return np.frombuffer(buf, dtype="<c16").reshape((num_g_corr, num_g_corr, nfs), order="F").copy(order="F")


def num_records(path: str, num_g_corr: int, nfs: int) -> int:
    return os.path.getsize(path) // (lrcoul(num_g_corr, nfs) * BYTE_REAL)

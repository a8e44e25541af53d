"""On-disk records of the reference -- W(q), Sigma_c(k, omega) and the wavefunction buffer --
so that the GPU path can exchange the screened Coulomb interaction, the self-energy and its input wavefunctions with an
unmodified ``gw.x``.  (The iotk *binary* XML container of sigma_io.f90 is not written: it needs the iotk library's
record framing; the direct-access records below carry the same arrays.)

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


# ---------------------------------------------------------------------------------------------------------------------
# Sigma_c(k, omega) records: opengwfil.f90:61-64 (lrsigma = 2 * num_g_corr**2 reals), written by
# davcio(sigma_root(:,:,ifreq), lrsigma, iunsigma, irec, 1) with irec = (ikpt - 1) * num_sigma + ifreq
# (phys/corr/src/sigma.f90:391-394) and read back by sigma_matel (sigma_expect_file).
def lrsigma(num_g_corr: int) -> int:
    return 2 * num_g_corr * num_g_corr


def _write_record(path: str, irec: int, a: np.ndarray) -> None:
    if irec < 1:
        raise ValueError("record index is 1-based")
    buf = np.asfortranarray(a).astype("<c16").tobytes(order="F")
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    with open(path, "r+b" if os.path.exists(path) else "w+b") as f:
        f.seek((irec - 1) * len(buf))
        f.write(buf)


def _read_record(path: str, irec: int, shape) -> np.ndarray:
    rec = int(np.prod(shape)) * 2 * BYTE_REAL
    with open(path, "rb") as f:
        f.seek((irec - 1) * rec)
        buf = f.read(rec)
    if len(buf) != rec:
        raise IOError(f"record {irec} of {path} is incomplete ({len(buf)} of {rec} bytes)")
    return np.frombuffer(buf, dtype="<c16").reshape(shape, order="F").copy(order="F")


def write_sigma_c(path: str, ikpt: int, sigma_c: np.ndarray) -> None:
    """All frequency records of one k-point: sigma_c(num_g_corr, num_g_corr, num_sigma) (sigma.f90:391-394)."""
    a = np.asarray(sigma_c, dtype=np.complex128)
    if a.ndim != 3 or a.shape[0] != a.shape[1]:
        raise ValueError("sigma_c must be (num_g_corr, num_g_corr, num_sigma)")
    for ifreq in range(a.shape[2]):
        _write_record(path, (ikpt - 1) * a.shape[2] + ifreq + 1, a[:, :, ifreq])


def read_sigma_c(path: str, ikpt: int, num_g_corr: int, num_sigma: int) -> np.ndarray:
    out = np.zeros((num_g_corr, num_g_corr, num_sigma), dtype=np.complex128, order="F")
    for ifreq in range(num_sigma):
        out[:, :, ifreq] = _read_record(path, (ikpt - 1) * num_sigma + ifreq + 1, (num_g_corr, num_g_corr))
    return out


# ---------------------------------------------------------------------------------------------------------------------
# Wavefunction buffer: algo/io/src/openfilq.f90:55 (lrwfc = nbnd * npwx * npol COMPLEX words per record, [QE] open_buffer
# -> direct-access file <tmp_dir_gw><prefix>.wfc), record ik = evc(npwx * npol, nbnd); read by get_buffer(evc, lrwfc, iuwfc, ik)
# at solve_linter.f90:318-330 and green.f90 / sigma_matel.f90:133.
def lrwfc(nbnd: int, npwx: int, npol: int = 1) -> int:
    return nbnd * npwx * npol


def write_wfc_record(path: str, ik: int, evc: np.ndarray) -> None:
    a = np.asarray(evc, dtype=np.complex128)
    if a.ndim != 2:
        raise ValueError("evc must be (npwx * npol, nbnd)")
    _write_record(path, ik, a)


def read_wfc_record(path: str, ik: int, npwx: int, nbnd: int, npol: int = 1) -> np.ndarray:
    return _read_record(path, ik, (npwx * npol, nbnd))

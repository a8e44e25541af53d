"""W(q) record format of the reference (opengwfil.f90:51-55, do_stern.f90:236, sigma.f90:306-331)."""
import numpy as np
import pytest

from sternheimergw_b200 import wfile


def test_record_layout_is_fortran_direct_access(tmp_path):
    ngc, nfs = 5, 3
    rng = np.random.default_rng(0)
    path = wfile.coul_filename(str(tmp_path), "si")
    assert path.endswith("_gw0/si.coul1")
    recs = {}
    for iq in (3, 1, 2):                                         # out-of-order writes, like images finishing at random
        a = rng.standard_normal((ngc, ngc, nfs)) + 1j * rng.standard_normal((ngc, ngc, nfs))
        recs[iq] = a
        wfile.write_w_record(path, iq, a)
    assert wfile.lrcoul(ngc, nfs) == 2 * ngc * ngc * nfs
    assert wfile.num_records(path, ngc, nfs) == 3
    raw = np.fromfile(path, dtype="<f8")
    assert raw.size == 3 * wfile.lrcoul(ngc, nfs)                 # no record markers
    # record 2 starts at (2-1)*lrcoul reals; element (ig, igp, iw) at 2*(ig + ngc*(igp + ngc*iw)) (column-major, re then im)
    off = wfile.lrcoul(ngc, nfs)
    ig, igp, iw = 3, 1, 2
    k = off + 2 * (ig + ngc * (igp + ngc * iw))
    assert raw[k] == recs[2][ig, igp, iw].real and raw[k + 1] == recs[2][ig, igp, iw].imag
    for iq, a in recs.items():
        assert np.array_equal(wfile.read_w_record(path, iq, ngc, nfs), a)
    with pytest.raises(IOError):
        wfile.read_w_record(path, 4, ngc, nfs)
    with pytest.raises(ValueError):
        wfile.write_w_record(path, 0, recs[1])


def test_roundtrip_of_an_inverted_epsilon_through_the_oracle(tmp_path):
    """do_stern.f90:220-236 on the host: unfold -> invert -> write; sigma.f90:331 reads the same numbers back."""
    import oracle
    ngc, nfs = 7, 2
    rng = np.random.default_rng(1)
    scr = np.asfortranarray(0.1 * (rng.standard_normal((ngc, nfs, ngc)) + 1j * rng.standard_normal((ngc, nfs, ngc))))
    for i in range(ngc):
        scr[i, :, i] += 2.0
    igu = np.arange(1, ngc + 1, dtype=np.int32)
    w, info = oracle.invert_epsilon(oracle.unfold_w(ngc, nfs, igu, scr), lgamma=False)
    assert info == 0
    path = wfile.coul_filename(str(tmp_path), "c")
    wfile.write_w_record(path, 2, w)
    back = wfile.read_w_record(path, 2, ngc, nfs)
    assert np.array_equal(back, w)
    assert np.count_nonzero(np.fromfile(path, dtype="<f8")[:wfile.lrcoul(ngc, nfs)]) == 0   # record 1 not written yet: zeros


def test_sigma_and_wfc_records_roundtrip(tmp_path):
    """Sigma_c records (sigma.f90:391-394: irec = (ikpt-1) num_sigma + ifreq, lrsigma = 2 ngc^2 reals) and the wavefunction
    buffer (openfilq.f90:55: lrwfc = nbnd npwx npol complex words)."""
    import numpy as np
    from sternheimergw_b200 import wfile
    rng = np.random.default_rng(0)
    ngc, nsig = 5, 3
    p = str(tmp_path / "_gw0" / "si.sigma1")
    s1 = rng.standard_normal((ngc, ngc, nsig)) + 1j * rng.standard_normal((ngc, ngc, nsig))
    s2 = rng.standard_normal((ngc, ngc, nsig)) + 1j * rng.standard_normal((ngc, ngc, nsig))
    wfile.write_sigma_c(p, 2, s2)            # out of order, like images finishing at different times
    wfile.write_sigma_c(p, 1, s1)
    assert np.array_equal(wfile.read_sigma_c(p, 1, ngc, nsig), s1)
    assert np.array_equal(wfile.read_sigma_c(p, 2, ngc, nsig), s2)
    import os
    assert os.path.getsize(p) == 2 * nsig * wfile.lrsigma(ngc) * 8
    # raw layout: record 4 (= k 2, frequency 1) holds s2[:, :, 0] column-major
    raw = np.fromfile(p, dtype="<c16")
    assert np.array_equal(raw[3 * ngc * ngc:4 * ngc * ngc], s2[:, :, 0].ravel(order="F"))
    w = str(tmp_path / "si.wfc")
    npwx, nbnd = 7, 4
    evc = rng.standard_normal((npwx, nbnd)) + 1j * rng.standard_normal((npwx, nbnd))
    wfile.write_wfc_record(w, 3, evc)
    assert np.array_equal(wfile.read_wfc_record(w, 3, npwx, nbnd), evc)
    assert os.path.getsize(w) == 3 * wfile.lrwfc(nbnd, npwx) * 16
    import pytest
    with pytest.raises(IOError):
        wfile.read_wfc_record(w, 4, npwx, nbnd)

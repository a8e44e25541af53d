"""Extract the known-answer vectors of the reference's AAA unit test (vendor/analytic/test/testAAA.pf) into
tests/golden/testAAA.npz, so that the oracle (oracle/sigma.py) and the GPU fit can be pinned to them on machines
where /root/reference does not exist.  Run once in the build container:  python tools/aaa_golden_extract.py
"""
from __future__ import annotations

import re
import sys
from pathlib import Path

import numpy as np

SRC = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference/vendor/analytic/test/testAAA.pf")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "testAAA.npz"

NUM = r"[-+]?\d+\.?\d*(?:[eE][-+]?\d+)?(?:_dp)?"


def _f(tok: str) -> float:
    return float(tok.replace("_dp", ""))


def bracket(text: str, start: int) -> str:
    """text of the array constructor [ ... ] that starts at or after `start` (continuation lines joined)."""
    i = text.index("[", start)
    j = text.index("]", i)
    return text[i + 1:j].replace("&", " ")


def parse_values(body: str) -> np.ndarray:
    if "CMPLX" in body:
        out = [complex(_f(a), _f(b)) for a, b in re.findall(rf"CMPLX\(\s*({NUM})\s*,\s*({NUM})\s*,\s*KIND=dp\)", body)]
        return np.array(out, dtype=complex)
    return np.array([_f(t) for t in re.findall(NUM, body)], dtype=float)


def grab(text: str, pattern: str, after: int = 0) -> np.ndarray:
    m = re.compile(pattern).search(text, after)
    assert m, pattern
    return parse_values(bracket(text, m.end() - 1))


def main():
    t = SRC.read_text()
    g = {}
    # make_realistic_example (testAAA.pf:75-160): 35 imaginary frequencies of a GW calculation, 14 support points
    g["real_zz"] = 1j * grab(t, r"example%zz = imag \* \[")
    g["real_ff"] = grab(t, r"example%ff = \[")
    g["real_selection"] = grab(t, r"example%selection = \[").astype(int)
    g["real_weight"] = grab(t, r"example%ref_weight = \[")
    g["real_pole"] = grab(t, r"example%ref_pole = \[")
    g["real_residual"] = grab(t, r"example%ref_residual = \[")
    g["real_threshold"] = np.array(1e-10)
    # test_tangent (testAAA.pf:221-270)
    p = t.index("SUBROUTINE test_tangent")
    g["tan_zz"] = 1j * grab(t, r"zz\(11\) = \[", p)
    g["tan_pos"] = 1j * grab(t, r"pos_ref\(6\) = \[", p)
    g["tan_val"] = 1j * grab(t, r"val_ref\(6\) = \[", p)
    g["tan_weight"] = grab(t, r"weight_ref\(6\) = \[", p)
    # test_evaluate_tangent (testAAA.pf:272-310)
    p = t.index("SUBROUTINE test_evaluate_tangent")
    g["tan_eval_zz"] = grab(t, r"zz\(11\) = \[", p) + 1j
    g["tan_eval_ff"] = grab(t, r"ff_ref\(11\) = \[", p)
    # test_threshold (testAAA.pf:312-352)
    p = t.index("SUBROUTINE test_threshold")
    g["thr_list"] = grab(t, r"thres_list\(6\) = \[", p)
    g["thr_steps"] = grab(t, r"step_list\(6\) = \[", p).astype(int)
    # test_pole_residual (testAAA.pf:387-422)
    p = t.index("SUBROUTINE test_pole_residual\n")
    num = lambda name: complex(*map(_f, re.search(rf"{name} = ({NUM}) \+ ({NUM}) \* imag", t[p:]).groups()))
    pole1, pole2, res1, res2 = num("pole1"), num("pole2"), num("res1"), num("res2")
    big = _f(re.search(rf"ref_pole\(5\) = \[({NUM}) \* imag", t[p:]).group(1))
    rbig = _f(re.search(rf"ref_res\(5\) = \[({NUM}) \+ c_zero", t[p:]).group(1))
    g["tan_pole"] = np.array([1j * big, pole1, -np.conj(pole1), pole2, -np.conj(pole2)])
    g["tan_res"] = np.array([rbig, res1, np.conj(res1), res2, np.conj(res2)])
    assert g["real_zz"].size == 35 and g["real_ff"].size == 35 and g["real_selection"].size == 14
    assert g["real_weight"].size == 14 and g["real_pole"].size == 13 and g["real_residual"].size == 13
    assert g["tan_zz"].size == 11 and g["tan_pos"].size == 6 and g["tan_eval_ff"].size == 11
    np.savez(OUT, **g)
    print("wrote", OUT, {k: v.shape for k, v in g.items()})


if __name__ == "__main__":
    main()

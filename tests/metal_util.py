"""Synthetic metallic stand-in shared by the oracle and GPU tests of the lgauss path."""
import numpy as np

import synth


def metal_system(nbnd_all=8, ef=None, degauss=0.05, ngauss=0, alpha_scale=1.0, nk=1, wg_one=False):
    s = synth.build_lattice("tiny", 10.26, synth.FCC, [[0.125] * 3, [-0.125] * 3], ["Si", "Si"], 6.0)
    syn = synth.attach_kpoints(s, synth.mp_grid(s.bg, nk), [0.5, 0.5, 0.5], nbnd=nbnd_all)
    if ef is None:
        ef = 0.62          # bands at 0.59 Ry (L) and 0.68 Ry (Gamma) straddle it: occupations 0.80 and 0.04 at degauss = 0.05
    for kp in syn.kpairs:
        kp.kq.alpha_pv *= alpha_scale
    syn.alpha_pv *= alpha_scale
    synth.make_metal(syn, ef, degauss, ngauss)
    if wg_one:
        for m in syn.metal.pairs:
            m.wg_over_wk[:] = 1.0
    return syn

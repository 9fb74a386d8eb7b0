"""Synthetic workloads of BASELINE.json.  Pure host-side Python: the
bath is the reference's own default (init_dmft_bath, ED_BATH/dmft_aux.f90:102-133, HWBAND=2),
restated here so that the product does not depend on oracle/."""
import math

import numpy as np


def init_dmft_bath(norb, nbath, nspin=1, hwband=2.0):
    """bath_type=normal default bath: e,v shaped (nspin, norb, nbath) [dmft_aux.f90:102-133]."""
    e = np.zeros((nspin, norb, nbath))
    v = np.zeros((nspin, norb, nbath))
    e[:, :, 0] = -hwband
    e[:, :, nbath - 1] = hwband
    nh = nbath // 2
    if nbath % 2 == 0 and nbath >= 4:
        de = hwband / max(nh - 1, 1)
        e[:, :, nh - 1] = -0.1
        e[:, :, nh] = 0.1
        for i in range(2, nh):
            e[:, :, i - 1] = -hwband + (i - 1) * de
            e[:, :, nbath - i] = hwband - (i - 1) * de
    elif nbath % 2 != 0 and nbath >= 3:
        de = hwband / nh
        e[:, :, nh] = 0.0
        for i in range(2, nh + 1):
            e[:, :, i - 1] = -hwband + (i - 1) * de
            e[:, :, nbath - i] = hwband - (i - 1) * de
    v[:, :, :] = max(0.1, 1.0 / math.sqrt(nbath))
    return e, v


def _single_band(nbath, n):
    e, v = init_dmft_bath(1, nbath)
    return dict(norb=1, nbath=nbath, nspin=1, uloc=(2.0,), ust=0.0, jh=0.0, jx=0.0, jp=0.0, xmu=0.0,
                hfmode=True, bath_e=e, bath_v=v, nup=n, ndw=n)


def config(name):
    """C1..C5 of BASELINE.json; returns kwargs for Oracle(...) / edgpu.Solver(...) plus nup, ndw."""
    name = name.upper()
    if name == "C1":
        return _single_band(7, 4)       # Ns=8,  4:4, dim 4 900
    if name == "C2":
        return _single_band(13, 7)      # Ns=14, 7:7, dim 11 778 624
    if name == "C3":
        return _single_band(15, 8)      # Ns=16, 8:8, dim 165 636 900
    if name == "C5":
        return _single_band(17, 9)      # Ns=18, 9:9, dim 2 363 904 400
    if name == "C4":                    # two-orbital Kanamori, Ns=10, 5:5
        e, v = init_dmft_bath(2, 4)     # e = -2,-0.1,0.1,2 ; V = 0.5
        return dict(norb=2, nbath=4, nspin=1, uloc=(2.0, 2.0), ust=1.0, jh=0.5, jx=0.5, jp=0.5, xmu=0.0,
                    hfmode=True, bath_e=e, bath_v=v, nup=5, ndw=5)
    if name in ("ORB2", "ORB2B", "ORB3"):  # density-density multi-orbital models for ed_total_ud = F (Jx = Jp = 0)
        norb, nbath = {"ORB2": (2, 2), "ORB2B": (2, 4), "ORB3": (3, 2)}[name]
        e, v = init_dmft_bath(norb, nbath)
        k = np.arange(nbath, dtype=np.float64).reshape(1, 1, nbath)
        o = np.arange(norb, dtype=np.float64).reshape(1, norb, 1)
        v = v * (1.0 + 0.2 * o) + 0.05 * k      # orbital- and level-dependent hybridisations
        e = e + 0.1 * o
        tot = norb * (nbath + 1) // 2
        return dict(norb=norb, nbath=nbath, nspin=1, uloc=tuple(2.0 + 0.5 * i for i in range(norb)), ust=1.0, jh=0.4, jx=0.0, jp=0.0,
                    xmu=0.1, hfmode=True, bath_e=e, bath_v=v, nup=tot, ndw=tot)
    if name.startswith("NS"):           # e.g. NS10 -> single band Ns=10 half filling; NS10V: level-dependent V_k
        vary = name.endswith("V")
        ns = int(name[2:-1] if vary else name[2:])
        cfg = _single_band(ns - 1, ns // 2)
        if vary:                        # a fitted bath: distinct hybridisations and asymmetric levels
            k = np.arange(ns - 1, dtype=np.float64)
            cfg["bath_v"] = (0.25 + 0.07 * k - 0.004 * k * k).reshape(1, 1, ns - 1)
            cfg["bath_e"] = cfg["bath_e"] + 0.013 * (k - 1.5).reshape(1, 1, ns - 1)
            cfg["xmu"] = 0.1
        return cfg
    raise KeyError(name)


def bench_vector(dim, offset=0, dtype=np.float64):
    """v_i = sin(0.37 i) + 0.1 (1-based i), un-normalised slice [offset, offset+dim)."""
    i = np.arange(offset + 1, offset + dim + 1, dtype=np.float64)
    return (np.sin(0.37 * i) + 0.1).astype(dtype)


def solver_kwargs(cfg):
    return {k: v for k, v in cfg.items() if k not in ("nup", "ndw")}

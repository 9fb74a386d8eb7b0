"""Generates the committed fixtures in tests/golden/ from the CPU oracle.

The reference has no golden vectors and cannot run here, so these files freeze
the ORACLE's outputs (they pin regressions; they are not reference outputs).  Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "dmft-lanc-ed_b200"))
import oracle as O  # noqa: E402
from edgpu import configs  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def c1():
    cfg = configs.config("C1")
    o = O.Oracle(**configs.solver_kwargs(cfg))
    with o.sector(4, 4) as s:
        v = configs.bench_vector(s.dim)
        v /= np.linalg.norm(v)
        e0, gs, a, b = s.lanc_eigh(v0=np.ones(s.dim) / np.sqrt(s.dim))
        rp, cols, vals = s.hup()
        out = dict(map_up=s.map_up(), hup_rowptr=rp, hup_cols=cols, hup_vals=vals, h0d=s.h0d(), hv=s.spmatvec(v),
                   e0=e0, alanc=a, blanc=b, gs=gs)
    r = o.build_gf_normal(4, 4, gs, e0, 1, lmats=64, lreal=64)
    out.update(gmats=r["gmats"], greal=r["greal"], wm=r["wm"], wr=r["wr"],
               add_norm2=r["chains"][0]["norm2"], add_alanc=r["chains"][0]["alanc"], add_blanc=r["chains"][0]["blanc"],
               rem_norm2=r["chains"][1]["norm2"], rem_alanc=r["chains"][1]["alanc"], rem_blanc=r["chains"][1]["blanc"])
    sig, _ = o.sigma_normal(1, 1, 1j * r["wm"], r["gmats"])
    out["smats"] = sig
    np.savez_compressed(os.path.join(HERE, "c1_golden.npz"), **out)


def c4():
    cfg = configs.config("C4")
    o = O.Oracle(**configs.solver_kwargs(cfg))
    with o.sector(5, 5) as s:
        v = configs.bench_vector(s.dim)
        v /= np.linalg.norm(v)
        e0, gs, a, b = s.lanc_eigh(v0=np.ones(s.dim) / np.sqrt(s.dim))
        rpn, coln, valn = s.hnd()
        out = dict(e0=e0, alanc=a, blanc=b, hv=s.spmatvec(v), hnd_rowptr=rpn, hnd_cols=coln, hnd_vals=valn)
    for iorb in (1, 2):
        r = o.build_gf_normal(5, 5, gs, e0, iorb, lmats=64, lreal=16)
        out["gmats_l%d" % iorb] = r["gmats"]
        out["add_alanc_l%d" % iorb] = r["chains"][0]["alanc"]
        out["add_blanc_l%d" % iorb] = r["chains"][0]["blanc"]
        out["add_norm2_l%d" % iorb] = r["chains"][0]["norm2"]
        out["wm"] = r["wm"]
    np.savez_compressed(os.path.join(HERE, "c4_golden.npz"), **out)


def round2():
    """Small fixtures of the round-2 oracle additions: susceptibility chains, the orbital-resolved operator, the phonon
    terms, hybrid and replica baths (tests/test_oracle_known_answers.py::test_golden_fixture_round2)."""
    out = {}
    cfg = configs.config("NS6")
    o = O.Oracle(**configs.solver_kwargs(cfg))
    with o.sector(3, 3) as s:
        e0, gs, _, _ = s.lanc_eigh(v0=np.ones(s.dim) / np.sqrt(s.dim))
        v = configs.bench_vector(s.dim * 3)
        v /= np.linalg.norm(v)
        out["ph_hv"] = s.spmatvec_ph(v, 2, (0.7,), 0.4)
        out["ph_e0"] = s.lanc_eigh_ph(2, (0.7,), 0.4, np.ones(s.dim * 3) / np.sqrt(s.dim * 3))[0]
    for kind in (0, 1):
        ch = o.chi_chain(3, 3, gs, kind, 1, 1, ngfiter=30)
        out["chi%d_norm2" % kind], out["chi%d_alanc" % kind], out["chi%d_blanc" % kind] = ch["norm2"], ch["alanc"], ch["blanc"]
    cfg2 = configs.config("ORB2")
    o2 = O.Oracle(**configs.solver_kwargs(cfg2))
    with o2.sector_orbs([1, 2], [2, 1]) as so:
        vo = configs.bench_vector(so.dim)
        vo /= np.linalg.norm(vo)
        out["orbs_isector"], out["orbs_h0d"], out["orbs_hv"] = so.isector, so.h0d(), so.spmatvec(vo)
        m, rp, cols, vals = so.factor(2)
        out["orbs_f2_map"], out["orbs_f2_rowptr"], out["orbs_f2_cols"], out["orbs_f2_vals"] = m, rp, cols, vals
        out["orbs_e0"] = so.lanc_eigh(v0=np.ones(so.dim) / np.sqrt(so.dim))[0]
    common = dict(norb=2, nspin=1, uloc=(2.0, 1.5), ust=0.8, jh=0.3, jx=0.3, jp=0.3, xmu=0.2, hfmode=True)
    hb = np.zeros((1, 1, 2, 2, 2))
    for kp in range(2):
        hb[0, 0, :, :, kp] = [[-0.8 + 0.7 * kp, 0.25 - 0.1 * kp], [0.25 - 0.1 * kp, 0.3 + 0.2 * kp]]
    orep = O.Oracle(nbath=2, bath_type=2, bath_v=np.array([[0.4, 0.7]]), bath_h=hb, **common)
    ohyb = O.Oracle(nbath=3, bath_type=1, bath_e=np.array([[[-1.0, 0.2, 0.9]]]),
                    bath_v=np.array([[[0.5, 0.3, 0.2], [0.1, 0.4, 0.6]]]), **common)
    for tag, oo, sec in (("rep", orep, (3, 3)), ("hyb", ohyb, (2, 3))):
        with oo.sector(*sec) as s:
            vv = configs.bench_vector(s.dim)
            vv /= np.linalg.norm(vv)
            rp, cols, vals = s.hup()
            out[tag + "_hup_rowptr"], out[tag + "_hup_cols"], out[tag + "_hup_vals"] = rp, cols, vals
            out[tag + "_h0d"], out[tag + "_hv"] = s.h0d(), s.spmatvec(vv)
    np.savez_compressed(os.path.join(HERE, "round2_golden.npz"), **out)


if __name__ == "__main__":
    round2()
    c1()
    c4()
    print("golden fixtures written")

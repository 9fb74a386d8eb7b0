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


if __name__ == "__main__":
    c1()
    c4()
    print("golden fixtures written")

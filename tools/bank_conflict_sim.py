"""How many shared-memory wavefronts do the gathers of k_fcol need, and how few could a smarter lane <-> row assignment
need?  CPU model of the LDS.64 bank behaviour on the real index data of a sector (the oracle's spH0ups CSR):
a warp handles 32 consecutive rows; per slot (s-th source of every row) the 32 lanes issue one 8-byte load each, which
the hardware serves as two half-warps; inside a half-warp two lanes conflict when their addresses differ but fall into
the same bank pair ((index mod 16) for doubles).  Wavefronts of a half-warp = the largest number of distinct addresses
in one bank pair.  Compared: (a) lanes = rows in order (what the kernel does; ncu counts 34.2 wavefronts per 32 rows on
C3), (b) the best split of the 32 rows of a warp into its two half-warps found by local search (global accesses stay
coalesced: same 32 rows) -- for three orders of the slots inside a row (CSR, bath bit ascending, bath bit descending).
usage: python tools/bank_conflict_sim.py [NS] [NUP]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "dmft-lanc-ed_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import oracle as O  # noqa: E402
from edgpu import configs  # noqa: E402


def half_cost(idx):
    """wavefronts of one half-warp instruction: idx = the 16 source indices (-1 = padding: the zero slot, one address)"""
    idx = idx[idx >= 0]
    if idx.size == 0:
        return 1
    u = np.unique(idx)
    return int(np.bincount(u & 15, minlength=16).max())


def warp_cost(src, rows_a, rows_b):
    return sum(half_cost(src[rows_a, s]) + half_cost(src[rows_b, s]) for s in range(src.shape[1]))


def main():
    ns = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    nup = int(sys.argv[2]) if len(sys.argv) > 2 else ns // 2
    cfg = configs.config("NS%d" % ns)
    o = O.Oracle(**configs.solver_kwargs(cfg))
    with o.sector(nup, 1) as s:                                    # the up factor only depends on Nup
        rp, cols, vals = s.hup()
        n = s.dimup
    w = int(np.diff(rp).max())
    src = -np.ones((n, w), np.int64)
    for i in range(n):
        k = rp[i + 1] - rp[i]
        src[i, :k] = cols[rp[i]:rp[i + 1]]
    # slot orders: CSR (ascending source index: the uniform-V kernel), by bath bit ascending (the k-ordered entries of
    # the fitted-bath kernel), by bath bit DESCENDING (candidate: the high bits, which are the same for all rows of a low
    # group, get the same slots in every row of the group, so a slot's sources are consecutive rows of ONE partner group)
    with o.sector(nup, 1) as s2:
        mp = s2.map_up()
    def by_bit(desc):
        out = -np.ones((n, w), np.int64)
        for i in range(n):
            ks = []
            for q in range(rp[i], rp[i + 1]):
                x = int(mp[i]) ^ int(mp[cols[q]])
                ks.append(((x & ~1).bit_length() - 1, int(cols[q])))
            ks.sort(reverse=desc)
            for t, (_, c) in enumerate(ks):
                out[i, t] = c
        return out
    orders = {"CSR order": src, "bath bit ascending": by_bit(False), "bath bit descending": by_bit(True)}
    rng = np.random.default_rng(0)
    tot = {}
    nwarps = 0
    for name, sx in orders.items():
        tot["in order, " + name] = 0
        tot["best split, " + name] = 0
        for r0 in range(0, n - 31, 32):
            blk = sx[r0:r0 + 32]
            a, b = np.arange(16), np.arange(16, 32)
            base = warp_cost(blk, a, b)
            tot["in order, " + name] += base
            best, ba, bb = base, a.copy(), b.copy()
            for _ in range(3):                                      # local search: swap pairs between the halves while it helps
                improved = False
                for i in rng.permutation(16):
                    for j in rng.permutation(16):
                        ba[i], bb[j] = bb[j], ba[i]
                        c = warp_cost(blk, ba, bb)
                        if c < best:
                            best, improved = c, True
                        else:
                            ba[i], bb[j] = bb[j], ba[i]
                if not improved:
                    break
            tot["best split, " + name] += best
            if name == "CSR order":
                nwarps += 1
    print("Ns = %d, Nup = %d: %d rows, %d slots, %d warp-rows of 32" % (ns, nup, n, w, nwarps))
    print("ideal (conflict free): %.1f wavefronts per 32 rows" % (2.0 * w))
    for k, v in tot.items():
        print("%-40s %.2f wavefronts per 32 rows" % (k, v / nwarps))


if __name__ == "__main__":
    main()

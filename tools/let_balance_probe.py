#!/usr/bin/env python
"""Diagnostic: the per-rank work of a LET evaluation, reproduced on ONE GPU.  `world` ranks run as threads (tests/loopback.py);
a lock serialises their let_evaluate / let_estr_evaluate calls, so vpmb200_fmm_times gives every rank's section times alone on
the device.  Prints per rank: owned particles, cells, M2L / P2P pairs, counted work, section times.

    python tools/let_balance_probe.py [particles] [world] [field] [steps]
`steps` > 0 first advances the field that many RK3 + dynamic-SFS steps on one engine (bench.py's scheme and dt)."""
import sys
import threading

sys.path.insert(0, ".")
import numpy as np
import torch

import flowunsteady_b200 as fb
from flowunsteady_b200 import fields
from flowunsteady_b200.dist import ShardedField, partition
from tests.loopback import run_ranks

n = int(sys.argv[1]) if len(sys.argv) > 1 else 5_000_000
world = int(sys.argv[2]) if len(sys.argv) > 2 else 8
kind = sys.argv[3] if len(sys.argv) > 3 else "rings"
x, g, s = {"rings": fields.vortex_rings, "random": fields.random_field}[kind](n)
P = fb.new_particles(x, fields.floor_gamma(g), s)
nsteps = int(sys.argv[4]) if len(sys.argv) > 4 else 0
if nsteps:
    dyn = fb.default_schemes(kernel="gaussianerf", integration="rungekutta3", relaxation="pedrizzetti", uj="fmm", sfs="dynamic",
                             alpha=0.999, force_positive=1, clippings=1)
    with fb.Engine(P.shape[0], schemes=dyn) as e0:
        e0.upload(P)
        for _ in range(nsteps):
            e0.nextstep(1.0e-3, (0.0, 0.0, 0.0), relax=True)
        P = e0.download(np.zeros_like(P))
    sg = P[:, 6]
    print(f"after {nsteps} steps: sigma min/median/max {sg.min():.4g} {np.median(sg):.4g} {sg.max():.4g}; "
          f"|Gamma| max {np.abs(P[:, 3:6]).max():.4g}; non-finite {int((~np.isfinite(P[:, :7])).sum())}")
    for q in range(4):
        m = ((P[:, 0] >= 0) == bool(q & 2)) & ((P[:, 1] >= 0) == bool(q & 1))
        print(f"  quadrant x{'+' if q & 2 else '-'} y{'+' if q & 1 else '-'}: n {int(m.sum())} sigma max {sg[m].max():.4g} mean {sg[m].mean():.6g}")
parts = partition(P.shape[0], world)
lock = threading.Lock()
kw = dict(uj="fmm", sfs="constant")


class Serial:
    """Engine proxy: the heavy phases of one rank at a time."""

    def __init__(self, eng):
        self._e = eng

    def __getattr__(self, name):
        f = getattr(self._e, name)
        if name in ("let_evaluate", "let_estr_evaluate", "let_build"):
            def g(*a, **k):
                with lock:
                    r = f(*a, **k)
                    self._e.synchronize()
                    return r
            return g
        return f


def body(rank, coll):
    torch.cuda.set_device(0)
    lo, hi = parts[rank]
    eng = fb.Engine(hi - lo + 8, device=0, schemes=fb.default_schemes(**kw))
    eng.upload(P[lo:hi].copy())
    sf = ShardedField(Serial(eng), max_local=hi - lo + 8, device="cuda:0", coll=coll, fmm="let")
    out = []
    for it in range(3):                     # iteration 0 cuts by count, the next ones by the counted work
        sf.uj(True, True, True)
        eng.synchronize()
        wk = int(sf._dev_view(eng.let_work(), 8 ** sf.let_level, "<i8", torch.int64).sum().item())   # (already all-reduced)
        out.append(dict(n_own=int(sum(sf._let["recv"])), tree=eng.fmm_stats(), ms=eng.fmm_times(), work_total=wk))
    eng.close()
    return out


res = run_ranks(world, body)
for it in range(3):
    print(f"--- evaluation {it} ({'count-based' if it == 0 else 'work-weighted'} cut)")
    for r in range(world):
        o = res[r][it]
        print(f"rank {r}: n_own {o['n_own']:8d} cells {o['tree']['cells']:7d} leaves {o['tree']['leaves']:7d} m2l {o['tree']['m2l_pairs']:9d} "
              f"p2p {o['tree']['p2p_pairs']:8d}  ms {o['ms']}")
    near = [res[r][it]["ms"]["l2p_near"] + res[r][it]["ms"]["estr_near"] for r in range(world)]
    print(f"near field + E_str: max {max(near):.2f} mean {sum(near) / world:.2f}  imbalance {max(near) / (sum(near) / world):.2f}")

#!/usr/bin/env python
"""FMM vs direct on the GPU: error and time (development aid)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import flowunsteady_b200 as fb
from flowunsteady_b200 import fields


def rel_l2(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def relmax(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())


def run(n, field, p, theta, ncrit, sfs=False, nzs=0):
    x, g, s = field(n)
    P = fb.new_particles(x, g, s)
    out = {}
    for uj in ("direct", "fmm"):
        sch = fb.default_schemes(uj=uj, fmm_p=p, fmm_theta=theta, fmm_ncrit=ncrit, fmm_nonzero_sigma=nzs)
        with fb.Engine(n, schemes=sch) as eng:
            eng.upload(P)
            eng.uj(True, True, sfs); eng.synchronize()
            t0 = time.perf_counter()
            eng.uj(True, True, sfs); eng.synchronize()
            dt = time.perf_counter() - t0
            out[uj] = (eng.download(np.zeros_like(P)), dt, eng.fmm_stats() if uj == "fmm" else None)
    D, F = out["direct"][0], out["fmm"][0]
    print(f"{field.__name__:14s} N={n:8d} p={p} theta={theta} ncrit={ncrit} nzs={nzs}: direct {out['direct'][1]*1e3:9.2f} ms  fmm {out['fmm'][1]*1e3:8.2f} ms | "
          f"U l2 {rel_l2(F[:,9:12], D[:,9:12]):.2e} max {relmax(F[:,9:12], D[:,9:12]):.2e} | J l2 {rel_l2(F[:,15:24], D[:,15:24]):.2e} max {relmax(F[:,15:24], D[:,15:24]):.2e}"
          + (f" | SFS l2 {rel_l2(F[:,39:42], D[:,39:42]):.2e}" if sfs else "") + f" | {out['fmm'][2]}")


if __name__ == "__main__":
    ns = [int(a) for a in sys.argv[1:]] or [20000, 200000]
    for n in ns:
        for field in (fields.vortex_rings, fields.random_field):
            for p, theta, nzs in ((4, 0.4, 0), (2, 0.4, 1), (4, 0.4, 1), (6, 0.4, 1), (4, 0.3, 1), (6, 0.3, 1)):
                run(n, field, p, theta, 50, sfs=(p == 4 and theta == 0.4), nzs=nzs)
        run(n, fields.vortex_rings, 4, 0.4, 128, nzs=1)

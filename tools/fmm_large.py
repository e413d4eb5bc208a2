#!/usr/bin/env python
"""UJ_fmm at large N: time per evaluation and error on 2048 sampled particles against the direct kernel (uj_probe)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import flowunsteady_b200 as fb
from flowunsteady_b200 import fields

for n in [int(a) for a in sys.argv[1:]] or [10_000_000]:
    x, g, s = fields.random_field(n)
    idx = np.random.default_rng(1234).choice(n, 2048, replace=False)
    with fb.Engine(n, schemes=fb.default_schemes(uj="fmm")) as eng:
        P = fb.new_particles(x, g, s)
        eng.upload(P)
        del P
        eng.uj(); eng.synchronize()
        t0 = time.perf_counter(); eng.uj(); eng.synchronize(); dt = time.perf_counter() - t0
        stats = eng.fmm_stats()
        Ud = eng.uj_probe(x[idx])
        out = np.zeros((n, 43)); eng.download(out, field_mask=1 << 5)
        err = np.linalg.norm(out[idx, 9:12] - Ud) / np.linalg.norm(Ud)
    print(f"random N={n}: UJ_fmm {dt*1e3:.1f} ms/evaluation, rel l2 err U vs direct (2048 samples) {err:.2e}, {stats}", flush=True)

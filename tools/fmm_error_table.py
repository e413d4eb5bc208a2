#!/usr/bin/env python
"""Error of UJ_fmm against the direct kernel, split into its parts, on the BASELINE configurations.

For every field: 2048 sampled particles; relative L2 error of U and J of
  total        p = 4, ncrit = 50, theta = 0.4, nonzero_sigma = false  vs direct gaussianerf   (the reference's defaults)
  truncation   the same expansions with the SINGULAR kernel everywhere vs direct singular      (pure expansion error)
  regularised  nonzero_sigma = true vs direct gaussianerf     (no singular far field inside 5 / 4 / 3 sigma of clearance between
               the closest points of two cells: vpmb200_schemes.fmm_nonzero_sigma = 1 / 4 / 3)
and the time of one evaluation in each mode.  total - truncation is what using the singular far field inside the regularised
range costs: a property of the reference's scheme (oracle/fmm_oracle.c reproduces it, tests/test_gpu_fmm.py).

    python tools/fmm_error_table.py > profiles/r02_fmm_error_table.md"""
import sys
import time

sys.path.insert(0, ".")
import numpy as np

import flowunsteady_b200 as fb
from flowunsteady_b200 import fields

CASES = [("rotor hover 200k (configs[1])", lambda: fields.rotor_wake(200_000, nfil=101, nsteps_per_rev=72)),
         ("vortex rings 1M (configs[2])", lambda: fields.vortex_rings(1_000_000)),
         ("vahana 5M (configs[3])", lambda: fields.vahana_wake(5_000_000)),
         ("random 2M (configs[4])", lambda: fields.random_field(2_000_000))]


def rel(a, b):
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


print("| field | mode | rel L2 err U | rel L2 err J | ms / evaluation | cells | M2L pairs | P2P pairs |")
print("|---|---|---|---|---|---|---|---|")
for name, gen in CASES:
    x, g, s = gen()
    g = fields.floor_gamma(g)
    n = x.shape[0]
    P = fb.new_particles(x, g, s)
    idx = np.sort(np.random.default_rng(1234).choice(n, 2048, replace=False))
    truth = {}
    with fb.Engine(n, schemes=fb.default_schemes(uj="direct")) as e:
        e.upload(P)
        for kernel in ("gaussianerf", "singular"):
            e.set_schemes(fb.default_schemes(uj="direct", kernel=kernel))
            truth[kernel] = e.uj_probe(x[idx], want_J=True)
    for mode, kw, kernel in (("total (reference defaults)", dict(fmm_nonzero_sigma=0), "gaussianerf"),
                             ("truncation only (singular kernel)", dict(fmm_nonzero_sigma=0), "singular"),
                             ("nonzero_sigma = true (5 sigma clearance)", dict(fmm_nonzero_sigma=1), "gaussianerf"),
                             ("nonzero_sigma = true, 4 sigma clearance", dict(fmm_nonzero_sigma=4), "gaussianerf"),
                             ("nonzero_sigma = true, 3 sigma clearance", dict(fmm_nonzero_sigma=3), "gaussianerf")):
        with fb.Engine(n, schemes=fb.default_schemes(uj="fmm", kernel=kernel, fmm_p=4, fmm_ncrit=50, fmm_theta=0.4, **kw)) as e:
            e.upload(P)
            e.uj(); e.synchronize()
            t0 = time.perf_counter()
            for _ in range(3):
                e.uj()
            e.synchronize()
            ms = (time.perf_counter() - t0) / 3 * 1e3
            out = e.download(np.zeros_like(P), field_mask=fb.engine.FM_U | fb.engine.FM_J)
            st = e.fmm_stats()
        Ud, Jd = truth[kernel]
        print(f"| {name} | {mode} | {rel(out[idx, 9:12], Ud):.2e} | {rel(out[idx, 15:24], Jd):.2e} | {ms:.1f} | {st['cells']} | "
              f"{st['m2l_pairs']} | {st['p2p_pairs']} |", flush=True)

#!/usr/bin/env python
"""A few RK3 + dynamic-SFS + pedrizzetti steps through UJ_fmm at N particles — target of the ncu launch list, and a
wall-clock/CUDA-event breakdown of where a step's time goes (evaluation vs the rest)."""
import sys
import time
sys.path.insert(0, ".")
import numpy as np
import flowunsteady_b200 as fb
from flowunsteady_b200 import fields

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "rings"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
x, g, s = fields.vortex_rings(n) if kind == "rings" else fields.rotor_wake(n, nfil=101, nsteps_per_rev=360, p_per_step=2)
sch = fb.default_schemes(uj="fmm", sfs="dynamic", alpha=0.999, force_positive=1, clippings=1)
with fb.Engine(x.shape[0], schemes=sch) as eng:
    eng.upload(fb.new_particles(x, g, s))
    for label, fn in (("uj", lambda: eng.uj(True, True, False)), ("uj+estr", lambda: eng.uj(True, True, True)),
                      ("step", lambda: eng.nextstep(1e-3, (0.0, 0.0, 0.0), True))):
        fn(); eng.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        eng.synchronize()
        print(f"{kind} N={x.shape[0]} {label}: {(time.perf_counter() - t0) / steps * 1e3:.2f} ms, launches so far {eng.launch_count}, "
              f"tree {eng.fmm_stats()}, nonfinite {eng.count_nonfinite()}", flush=True)
    P = eng.download(np.zeros((x.shape[0], 43)))
    print("sigma range", P[:, 6].min(), P[:, 6].max(), "|Gamma| max", np.abs(P[:, 3:6]).max(), "C range", P[:, 36].min(), P[:, 36].max())

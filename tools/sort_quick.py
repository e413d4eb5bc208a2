#!/usr/bin/env python
"""Direct path with / without internal Morton ordering (development aid)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
import flowunsteady_b200 as fb
from flowunsteady_b200 import fields

for n in [int(a) for a in sys.argv[1:]] or [200000]:
    for field in (fields.vortex_rings, fields.random_field, fields.rotor_wake):
        x, g, s = field(n)
        P = fb.new_particles(x, g, s)
        res = {}
        for srt in (0, 1):
            with fb.Engine(P.shape[0], schemes=fb.default_schemes()) as eng:
                eng.set_option("direct_sort", srt)
                eng.upload(P)
                eng.uj(True, True, True); eng.synchronize()
                t0 = time.perf_counter(); eng.uj(); eng.synchronize(); t1 = time.perf_counter()
                eng.uj(True, True, True); eng.synchronize(); t2 = time.perf_counter()
                res[srt] = (eng.download(np.zeros_like(P)), t1 - t0, t2 - t1)
        d = np.abs(res[1][0][:, 9:24] - res[0][0][:, 9:24]).max() / np.abs(res[0][0][:, 9:24]).max()
        e = np.abs(res[1][0][:, 39:42] - res[0][0][:, 39:42]).max() / np.abs(res[0][0][:, 39:42]).max()
        print(f"{field.__name__:14s} N={P.shape[0]:8d}: uj unsorted {res[0][1]*1e3:9.2f} ms sorted {res[1][1]*1e3:9.2f} ms | uj+estr unsorted {res[0][2]*1e3:9.2f} sorted {res[1][2]*1e3:9.2f} ms | max rel diff UJ {d:.1e} SFS {e:.1e}")

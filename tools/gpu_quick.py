#!/usr/bin/env python
"""Quick GPU timing of one UJ evaluation at a few sizes + the FP64 peak (development aid; not the bench)."""
import ctypes as C
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import flowunsteady_b200 as fb
from flowunsteady_b200 import _lib, fields

L = _lib.lib()
tf, ms = C.c_double(), C.c_double()
L.vpmb200_measure_fp64_peak(0, 2000, 5, C.byref(tf), C.byref(ms))
print(f"FP64 DFMA peak: {tf.value:.2f} TFLOP/s ({ms.value:.2f} ms)")

sizes = [int(a) for a in sys.argv[1:]] or [50_000, 200_000]
for bits in (64, 32):
    for kernel in ("gaussianerf", "winckelmans", "singular"):
        for n in sizes:
            x, g, s = fields.vortex_rings(n)
            P = fb.new_particles(x, g, s)
            with fb.Engine(n, float_bits=bits, schemes=fb.default_schemes(kernel=kernel)) as eng:
                eng.upload(P)
                eng.uj(); eng.synchronize()
                t0 = time.perf_counter()
                reps = 3
                for _ in range(reps):
                    eng.uj()
                eng.synchronize()
                dt = (time.perf_counter() - t0) / reps
            print(f"fp{bits} {kernel:12s} N={n:8d}  {dt*1e3:9.2f} ms  {n*n/dt/1e9:8.1f} G inter/s  "
                  f"alg {n*n/dt*86/1e12:6.2f} TFLOP/s")
for n in sizes:
    x, g, s = fields.random_field(n)
    P = fb.new_particles(x, g, s)
    with fb.Engine(n, schemes=fb.default_schemes(kernel="gaussianerf", sfs="dynamic")) as eng:
        eng.upload(P)
        eng.uj(sfs=True, reset_sfs=True); eng.synchronize()
        t0 = time.perf_counter(); eng.uj(); eng.synchronize(); t1 = time.perf_counter()
        eng.uj(sfs=True, reset_sfs=True); eng.synchronize(); t2 = time.perf_counter()
        eng.nextstep(1e-3); eng.synchronize(); t3 = time.perf_counter()
    print(f"random N={n}: uj {1e3*(t1-t0):.2f} ms, uj+estr {1e3*(t2-t1):.2f} ms, RK3+dynSFS+relax step {1e3*(t3-t2):.2f} ms")

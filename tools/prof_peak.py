#!/usr/bin/env python
"""Runs the FP64 peak microbenchmark once (ncu capture target: how busy is the FP64 pipe in dfma_peak_kernel itself?)."""
import ctypes as C
import sys
sys.path.insert(0, ".")
from flowunsteady_b200 import _lib

o = (C.c_double * 8)()
rc = _lib.lib().vpmb200_measure_fp64_peak2(0, 2000, 1, o)
print(rc, dict(zip(("dfma_tflops", "ms", "sm_mhz", "pipe_tflops", "frac", "shape", "sms", "nominal_mhz"), list(o))))

#!/usr/bin/env python
"""A few UJ_fmm evaluations at N particles — target of the ncu launch list."""
import sys
sys.path.insert(0, ".")
import flowunsteady_b200 as fb
from flowunsteady_b200 import fields

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
nzs = int(sys.argv[2]) if len(sys.argv) > 2 else 0
copies = int(sys.argv[3]) if len(sys.argv) > 3 else 8
x, g, s = fields.vortex_rings(n)
with fb.Engine(n, schemes=fb.default_schemes(uj="fmm", fmm_nonzero_sigma=nzs)) as eng:
    eng.set_option("fmm_table_copies", copies)
    eng.upload(fb.new_particles(x, g, s))
    for _ in range(2):
        eng.uj(True, True, True)
    eng.synchronize()
    print(eng.fmm_stats())

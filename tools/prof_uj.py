#!/usr/bin/env python
"""One UJ evaluation on a single full wave of CTAs (296 x 256 targets) — the ncu capture target."""
import sys
sys.path.insert(0, ".")
import flowunsteady_b200 as fb
from flowunsteady_b200 import fields

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296 * 256
kernel = sys.argv[2] if len(sys.argv) > 2 else "gaussianerf"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
sfs = len(sys.argv) > 4 and sys.argv[4] == "sfs"     # also run the E_str pass (K2)
x, g, s = fields.vortex_rings(n)
with fb.Engine(n, schemes=fb.default_schemes(kernel=kernel)) as eng:
    eng.upload(fb.new_particles(x, g, s))
    for _ in range(reps):
        eng.uj(True, sfs, sfs)
    eng.synchronize()

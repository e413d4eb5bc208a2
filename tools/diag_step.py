#!/usr/bin/env python
"""Per-step wall time of the UJ_fmm dynamic-SFS step, engine-resident vs through vpm.ParticleField (host buffers)."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
import flowunsteady_b200 as fb
from flowunsteady_b200 import fields, vpm

n = 1_000_000
x, g, s = fields.vortex_rings(n)
P0 = fb.new_particles(x, g, s)
sch = fb.default_schemes(uj="fmm", fmm_p=4, fmm_ncrit=50, fmm_theta=0.4, fmm_nonzero_sigma=0, sfs="dynamic", alpha=0.999,
                         force_positive=1, clippings=1)
with fb.Engine(n, schemes=sch) as ef:
    ef.upload(P0)
    for k in range(4):
        t0 = time.perf_counter(); ef.uj(); ef.synchronize(); print("uj", k, (time.perf_counter() - t0) * 1e3, flush=True)
    ef.upload(P0)
    for k in range(8):
        l0 = ef.launch_count
        t0 = time.perf_counter(); ef.nextstep(1e-3, (0.0, 0.0, 0.0), True); ef.synchronize()
        print("engine step", k, (time.perf_counter() - t0) * 1e3, "launches", ef.launch_count - l0, ef.fmm_stats(), flush=True)
pf = vpm.ParticleField(n, formulation=vpm.rVPM, kernel=vpm.gaussianerf, UJ=vpm.UJ_fmm, SFS=vpm.SFS_Cd_twolevel_nobackscatter,
                       integration=vpm.rungekutta3, relaxation=vpm.pedrizzetti, sync="always", pinned=True)
pf.particles[:n] = P0
pf.np = n
for k in range(8):
    t0 = time.perf_counter(); vpm.nextstep(pf, 1e-3, relax=True); pf.engine.synchronize()
    print("host-buffer step", k, (time.perf_counter() - t0) * 1e3, pf.engine.fmm_stats(), flush=True)

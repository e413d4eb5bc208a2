"""FLOWUnsteady's wake-treatment runtime functions (/root/reference/src/FLOWUnsteady_processing.jl:50-187), running as a
device-side compaction (vpmb200_remove_where).  Same factory names, same arguments, same `extra_runtime_function`
signature `(sim, pfield, t, dt) -> Bool`; the surviving particles are left in the order the reference's loop produces.
`sim` needs `.nt` and `.vehicle.system.O` exactly as in the reference.
"""
from __future__ import annotations

import math

import numpy as np

from .engine import Engine


def remove_particles_strength(minGamma2: float, maxGamma2: float, every_nsteps: int = 1):
    def wake_treatment(sim, PFIELD, T, DT, *args, **optargs):
        if sim.nt % every_nsteps == 0:
            PFIELD.remove_where(Engine.REMOVE_STRENGTH, [minGamma2, maxGamma2])
        return False
    return wake_treatment


def remove_particles_lowstrength(crit_Gamma2: float, step: int):
    return remove_particles_strength(crit_Gamma2, math.inf, every_nsteps=step)


def remove_particles_sigma(minsigma: float, maxsigma: float, every_nsteps: int = 1):
    def wake_treatment(sim, PFIELD, T, DT, *args, **optargs):
        if sim.nt % every_nsteps == 0:
            PFIELD.remove_where(Engine.REMOVE_SIGMA, [minsigma, maxsigma])
        return False
    return wake_treatment


def remove_particles_box(Pmin, Pmax, step: int):
    def wake_treatment(sim, PFIELD, T, DT, *args, **optargs):
        if sim.nt % step == 0:
            O = np.asarray(sim.vehicle.system.O, dtype=np.float64)
            PFIELD.remove_where(Engine.REMOVE_BOX, list(Pmin) + list(Pmax) + list(O))
        return False
    return wake_treatment


def remove_particles_sphere(Rsphere2: float, step: int, Xoff=(0.0, 0.0, 0.0)):
    def wake_treatment(sim, PFIELD, T, DT, *args, **optargs):
        # NOTE the reference ignores `step` in this one (processing.jl:163-185): it runs every step
        O = np.asarray(sim.vehicle.system.O, dtype=np.float64) + np.asarray(Xoff, dtype=np.float64)
        PFIELD.remove_where(Engine.REMOVE_SPHERE, [Rsphere2] + list(O))
        return False
    return wake_treatment

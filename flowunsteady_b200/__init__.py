"""flowunsteady_b200 — B200-native rVPM particle-field engine behind FLOWVPM's ParticleField plugin API.

The compute path is hand-written sm_100a CUDA behind a C ABI (include/vpmb200.h -> libvpmb200.so); this
package is the host-side mirror of the FLOWVPM interface FLOWUnsteady drives (`flowunsteady_b200.vpm`), a thin
ctypes wrapper (`flowunsteady_b200.engine`) and synthetic workloads (`flowunsteady_b200.fields`).
"""
from . import _lib  # noqa: F401
from .engine import Engine, EngineError, MultiEngine, default_schemes, new_particles  # noqa: F401

__version__ = "0.1.0"

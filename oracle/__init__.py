"""CPU oracle for the rVPM hot path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  The product path (flowunsteady_b200) never does.  See vpm_oracle.h: parity unpinned.
"""
from .oracle import *  # noqa: F401,F403

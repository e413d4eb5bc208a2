#!/usr/bin/env python
"""Generates tests/golden/step_oracle.npz: particle matrices after two vpm.nextstep calls computed by the CPU oracle for a
fixed 96-particle field under every scheme family.  It freezes the ORACLE (a regression pin between rounds; the hot path
has no reference-side golden vectors — parity unpinned); the GPU tests compare against the same file.

Run:  python tools/gen_golden_step.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle import oracle as o  # noqa: E402

CASES = {
    "rk3_rvpm_pedrizzetti": dict(integration="rungekutta3"),
    "euler_cvpm_classic_corrected": dict(integration="euler", g=0.0, transposed=0, relaxation="correctedpedrizzetti"),
    "rk3_dynamic_sfs_twolevel": dict(integration="rungekutta3", sfs="dynamic", alpha=0.999, force_positive=1, clippings=1),
    "euler_dynamic_threelevel_controls": dict(integration="euler", sfs="dynamic", alpha=0.667, clippings=1, controls=3),
    "rk3_constant_sfs_winckelmans": dict(integration="rungekutta3", kernel="winckelmans", sfs="constant", clippings=1),
    "rk3_corespreading_reset": dict(integration="rungekutta3", viscous="corespreading", nu=2e-3, cs_sgm0=0.2, cs_beta=1.05,
                                    cs_itmax=5, cs_tol=1e-9),
}


def field():
    rng = np.random.default_rng(20261017)
    n = 96
    x = rng.random((n, 3))
    g = rng.standard_normal((n, 3)) * 0.5
    s = np.full(n, 0.2) * (0.8 + 0.4 * rng.random(n))
    static = (rng.random(n) < 0.1).astype(float)
    return o.new_field(x, g, s, static=static)


def main():
    out = {"P0": field()}
    for name, kw in CASES.items():
        P = field()
        if "corespreading" in name:
            P[:, o.SIGMA] = 0.2
        t, nt = 0.0, 0
        for _ in range(2):
            t, nt = o.nextstep(P, o.default_schemes(**kw), 5e-3, (1.0, -0.5, 0.25), relax=True, t=t, nt=nt)
        assert np.all(np.isfinite(P)), name
        out[name] = P
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "step_oracle.npz")
    np.savez_compressed(path, **out)
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    main()

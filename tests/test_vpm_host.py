"""CPU tests of the host-side FLOWVPM conveniences in flowunsteady_b200/vpm.py (monitors with log files, verbose helpers,
settings, run_vpm!'s loop).  They touch a field only through `particles`, `np`, `nt`, `t`, `monitors()` and `pull()`, so a
stand-in object drives them here; the GPU tests cover the same functions on a real ParticleField."""
import json
import os
import types

import numpy as np
import pytest

from flowunsteady_b200 import vpm


def _stub(n=50, seed=0):
    rng = np.random.default_rng(seed)
    P = np.zeros((n, 43))
    P[:, 36] = np.where(rng.random(n) < 0.3, 0.0, rng.random(n))
    f = types.SimpleNamespace(particles=P, np=n, nt=0, t=0.0, _dev_dirty=0, pulled=[])
    f.monitors = lambda: {"enstrophy": 1.25 + f.nt}
    f.pull = lambda mask=None: f.pulled.append(mask)
    return f


def test_cd_statistics_against_scipy():
    from scipy import stats
    rng = np.random.default_rng(3)
    C = np.where(rng.random(400) < 0.25, 0.0, rng.gamma(2.0, 0.1, 400))
    r0, mean, std, skew, kurt, cmin, cmax = vpm.cd_statistics(C)
    nz = C[C != 0]
    assert r0 == pytest.approx(1 - nz.size / C.size)
    assert mean == pytest.approx(nz.mean()) and std == pytest.approx(nz.std(ddof=1))
    assert skew == pytest.approx(stats.skew(nz)) and kurt == pytest.approx(stats.kurtosis(nz, fisher=False))
    assert (cmin, cmax) == (nz.min(), nz.max())
    assert vpm.cd_statistics(np.zeros(5)) == (1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
    assert vpm.cd_statistics(np.array([])) == (0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)


def test_monitors_append_to_out_and_to_csv(tmp_path):
    f = _stub()
    ens, cds = [], []
    for step in range(3):
        f.nt, f.t = step, 0.1 * step
        assert vpm.monitor_enstrophy(f, f.t, 0.1, save_path=str(tmp_path), run_name="run_", out=ens) is False
        assert vpm.monitor_Cd(f, f.t, 0.1, save_path=str(tmp_path), run_name="run_", out=cds) is False
    assert ens == [1.25, 2.25, 3.25]
    # the reference unpacks `t, rationzero, mean, stddev, skew, kurt, minC, maxC = out[end]` (src/FLOWUnsteady_monitors.jl:702)
    assert len(cds) == 3 and len(cds[-1]) == 8 and cds[-1][0] == pytest.approx(0.2)
    rows = open(tmp_path / "run_enstrophy.log").read().strip().split("\n")
    assert rows[0].startswith("nt,t (s),enstrophy") and len(rows) == 4 and rows[3].split(",")[0] == "2"
    rows = open(tmp_path / "run_Chistory.log").read().strip().split("\n")
    assert len(rows) == 4 and len(rows[1].split(",")) == 9
    f._dev_dirty = 1 << 10                       # lazy field: C is newer on the device -> monitor_Cd pulls it first
    vpm.monitor_Cd(f, 0.3, 0.1, out=cds)
    assert f.pulled == [1 << 10]
    f.np = 0
    assert vpm.monitor_enstrophy(f, 0.0, 0.1, out=ens) is False and len(ens) == 3


def test_verbose_helpers_and_create_path(tmp_path, capsys):
    sp = tmp_path / "out"
    vpm.create_path(str(sp), prompt=False)
    (sp / "stale.txt").write_text("x")
    vpm.create_path(str(sp), prompt=False)       # an existing directory is emptied
    assert sp.is_dir() and not (sp / "stale.txt").exists()
    f = types.SimpleNamespace(maxparticles=100)
    line1, line2, run_id, file_verbose, vprintln, t0 = vpm.initialize_verbose(True, str(sp), "case", f, 1e-3, 10)
    vprintln("hello", 1)
    vpm.finalize_verbose(t0, line1, vprintln, run_id, 0)
    log = open(file_verbose).read()
    assert "START" in log and "\thello" in log and "ELAPSED TIME" in log and "hello" in capsys.readouterr().out


def test_save_settings_roundtrip(tmp_path):
    f = types.SimpleNamespace(maxparticles=1000, R=np.float64, formulation=vpm.rVPM, kernel=vpm.gaussianerf,
                              viscous=vpm.CoreSpreading(1e-5, 0.02), UJ=vpm.UJ_fmm, integration=vpm.rungekutta3, transposed=True,
                              relaxation=vpm.pedrizzetti, fmm=vpm.FMM(p=4, ncrit=50, theta=0.4),
                              SFS=vpm.SFS_Cd_twolevel_nobackscatter)
    name = vpm.save_settings(f, "case", path=str(tmp_path))
    d = json.load(open(name))
    assert d["kernel"] == "gaussianerf" and d["formulation"] == {"f": 0.0, "g": 0.2} and d["fmm"]["ncrit"] == 50
    assert d["SFS"]["alpha"] == 0.999 and d["SFS"]["clippings"] == ["clipping_backscatter"]
    assert d["viscous"] == "CoreSpreading" and d["viscous_parameters"]["sgm0"] == 0.02
    assert os.path.isdir(vpm.utilities_path)


def test_particle_strength_exchange_is_refused_loudly():
    with pytest.raises(NotImplementedError):
        vpm.ParticleField(10, viscous=vpm.ParticleStrengthExchange(1e-5))


def test_run_vpm_loop_order_statics_and_saves(tmp_path):
    """vpm.run_vpm!'s loop on a stand-in field: statics are appended before nextstep and removed after it, the runtime
    function sees every step and can stop the run, and a restart file is written every nsteps_save steps."""
    calls = []

    class Field:
        maxparticles, R = 64, np.float64
        formulation, kernel, viscous = vpm.rVPM, vpm.gaussianerf, vpm.Inviscid()
        UJ, transposed, relaxation, fmm, SFS = vpm.UJ_direct, True, vpm.pedrizzetti, vpm.FMM(), vpm.SFS_none

        def __init__(self):
            self.particles = np.zeros((64, 43))
            self.particles[:4, 0] = np.arange(4)
            self.particles[:4, 3:6] = 1.0
            self.particles[:4, 6] = 0.1
            self.np, self.nt, self.t, self._dev_dirty = 4, 0, 0.0, 0

        def integration(self, pf, dt, relax=False, custom_UJ=None):
            calls.append(("step", pf.np, relax))
            pf.particles[:pf.np, 0] += dt

        def mark_dirty(self):
            pass

    def statics(pf, t, dt):
        pf.particles[pf.np, :] = 0.0
        pf.particles[pf.np, 42] = 1.0
        pf.np += 1

    seen = []

    def runtime(pf, t, dt):
        seen.append((pf.nt, pf.np))
        return pf.nt == 5

    f = Field()
    real_remove = vpm.remove_particle
    try:
        vpm.remove_particle = lambda pf, i: setattr(pf, "np", pf.np - 1) if i == pf.np - 1 else (_ for _ in ()).throw(AssertionError(i))
        vpm.run_vpm_(f, 0.5, 8, runtime_function=runtime, static_particles_function=statics, save_path=str(tmp_path / "run"),
                     run_name="pf", nsteps_save=2, verbose=False, prompt=False)
    finally:
        vpm.remove_particle = real_remove
    assert [c[1] for c in calls] == [5] * 5 and all(c[2] for c in calls)        # 4 particles + 1 static inside every step
    assert seen == [(k, 4) for k in range(6)]                                     # statics gone again; stopped at nt = 5
    assert f.t == pytest.approx(2.5) and f.particles[0, 0] == pytest.approx(2.5)
    files = sorted(os.listdir(tmp_path / "run"))
    assert "pf_settings.json" in files and "pf.log" in files
    assert [n for n in files if n.endswith(".h5")] == ["pf.0.h5", "pf.2.h5", "pf.4.h5", "pf.5.h5"]
    g = Field()
    vpm.read_(g, "pf.5.h5", path=str(tmp_path / "run"))
    assert g.np == 4 and g.nt == 5 and g.particles[0, 0] == pytest.approx(2.5)

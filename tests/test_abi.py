"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/vpmb200.h declares, and the
product path fails loudly (no CPU fallback) when no B200 is present."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "vpmb200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vpmb200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from flowunsteady_b200 import _lib
    L = C.CDLL(_lib.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/vpmb200.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes prototype in flowunsteady_b200/_lib.py"
    assert set(_lib.SYMBOLS) <= set(names)


def test_struct_layout_matches_header():
    from flowunsteady_b200 import _lib
    from oracle import oracle as o
    # first 22 members are shared with the oracle's struct (same order, same types)
    assert [f[0] for f in _lib.Schemes._fields_[:22]] == [f[0] for f in o.Schemes._fields_]
    s = _lib.Schemes()
    assert _lib.lib().vpmb200_default_schemes(C.byref(s)) == 0
    assert (s.kernel, s.f, s.g, s.transposed, s.relaxation, s.rlxf, s.integration) == (0, 0.0, 0.2, 1, 1, 0.3, 1)
    assert (s.fmm_p, s.fmm_ncrit, s.fmm_theta) == (4, 50, 0.4)   # vpm.FMM defaults, simulation.jl:43
    assert _lib.lib().vpmb200_tile_doubles() == 2570 and _lib.lib().vpmb200_tiles_for(257) == 2


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import flowunsteady_b200 as fb
    from flowunsteady_b200 import vpm
    with pytest.raises(fb.EngineError) as ei:
        fb.Engine(100)
    assert ei.value.code == -2            # VPMB200_ENODEVICE
    with pytest.raises(fb.EngineError):
        vpm.ParticleField(100)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "flowunsteady_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"#include\s+[\"<][^\n]*oracle", r"libvpm_oracle", r"\bvpmo_\w+\s*\("):
                    assert not re.search(pat, txt, flags=re.M), f"{f} reaches into oracle/ ({pat})"


def test_vpm_mirror_surface():
    """Every FLOWVPM name FLOWUnsteady uses exists in the mirror."""
    from flowunsteady_b200 import vpm
    names = """ParticleField nextstep add_particle remove_particle get_np get_particle iterate iterator get_X get_Gamma
        get_sigma get_vol get_circulation get_C get_U _reset_particles STATIC_INDEX SIGMA_INDEX rVPM cVPM formulation_rVPM
        formulation_cVPM gaussianerf winckelmans Kernel UJ_fmm UJ_direct FMM euler rungekutta3 pedrizzetti
        correctedpedrizzetti norelaxation relaxation_none Inviscid CoreSpreading zeta_fmm isinviscid iscorespreading
        _kernel_compatibility SFS_none SFS_Cs_nobackscatter SFS_Cd_twolevel_nobackscatter SFS_Cd_threelevel_nobackscatter
        DynamicSFS ConstantSFS Estr_fmm Estr_direct pseudo3level pseudo3level_positive clipping_backscatter
        control_directional control_magnitude control_sigmasensor isSFSenabled save read zeta_direct
        ParticleStrengthExchange monitor_enstrophy monitor_Cd save_settings create_path initialize_verbose
        finalize_verbose utilities_path run_vpm_""".split()
    for n in names:
        assert hasattr(vpm, n), n
    g, dg = vpm.gaussianerf.g_dgdr(1.3)
    from oracle import oracle as o
    assert (g, dg) == pytest.approx(o.g_dgdr("gaussianerf", 1.3), rel=1e-14)
    assert vpm.winckelmans.zeta(0.7) == pytest.approx(o.zeta("winckelmans", 0.7), rel=1e-14)
    assert vpm.SFS_Cd_twolevel_nobackscatter.alpha == 0.999 and vpm.SFS_Cd_threelevel_nobackscatter.alpha == 0.667
    assert vpm.gaussianerf in vpm._kernel_compatibility(vpm.CoreSpreading(1e-5, 0.1)) and \
        vpm.winckelmans not in vpm._kernel_compatibility(vpm.CoreSpreading(1e-5, 0.1))


def test_fields_generators():
    from flowunsteady_b200 import fields
    x, g, s = fields.vortex_rings(20_000)
    assert x.shape == (20_000, 3) and g.shape == (20_000, 3) and s.shape == (20_000,)
    # two rings of circulation 1: total |Gamma| ~ 2 * 2 pi R
    assert np.linalg.norm(g, axis=1).sum() == pytest.approx(2 * 2 * np.pi, rel=0.02)
    assert np.all(np.linalg.norm(g, axis=1) > 0)
    x, g, s = fields.random_field(1000)
    assert s[0] == pytest.approx(2.125 * 1000 ** (-1 / 3))
    x, g, s = fields.wing_wake(rows=100)
    assert x.shape[0] == 10_100 and s[0] == pytest.approx(0.0684, rel=2e-3)   # examples/wing: sigma = 0.0684 m
    x, g, s = fields.rotor_wake(70_000)
    assert s[0] == pytest.approx(0.01113, rel=2e-3)                            # rotorhover mid-low sigma

"""CPU tests of the synthetic workload generators (flowunsteady_b200/fields.py) that stand in for the BASELINE.json
configurations: particle counts and core sizes must be the reference examples' (citations in fields.py)."""
import numpy as np

from flowunsteady_b200 import fields


def test_counts_and_core_sizes_follow_the_reference_examples():
    x, g, s = fields.wing_wake(rows=100)
    assert x.shape == (10_100, 3) and abs(s[0] - 0.0684) < 1e-3            # examples/wing/wing.jl:46,58 ; sigma = lambda V dt
    x, g, s = fields.rotor_wake(70_000, nfil=41, nsteps_per_rev=36, p_per_step=4)
    assert abs(s[0] - 2.125 * 2 * np.pi * 0.12 / (36 * 4)) < 1e-12          # examples/rotorhover/rotorhover.jl:155-157
    assert 69_000 <= x.shape[0] <= 70_000
    x, g, s = fields.vortex_rings(100_000)
    assert x.shape[0] == 100_000
    x, g, s = fields.vahana_wake(60_000)
    assert x.shape == (60_000, 3) and s[0] == 0.0366                        # examples/vahana/vahana.jl:109
    c = np.array([-0.2 * 5.86, 0.0, 0.1 * 5.86])
    assert np.linalg.norm(x - c, axis=1).max() < 1.25 * 5.86                # remove_particles_sphere, vahana.jl:356
    x, g, s = fields.random_field(1000)
    assert np.allclose(s, 2.125 * 1000 ** (-1 / 3))


def test_no_generator_emits_a_zero_strength_component():
    """FLOWUnsteady floors every Gamma component to 5 eps when it sheds a particle (src/FLOWUnsteady_simulation.jl:464-468);
    a zero-strength particle would make the integrator's 1/|Gamma|^2 a 0/0 (that is the reference's behaviour too)."""
    tiny = 5 * np.finfo(np.float64).eps
    for x, g, s in (fields.wing_wake(rows=20), fields.vahana_wake(30_000), fields.vortex_rings(20_000),
                    fields.rotor_wake(20_000)):
        assert np.all(np.isfinite(x)) and np.all(np.isfinite(g)) and np.all(s > 0)
        assert np.all(np.linalg.norm(g, axis=1) >= tiny)
    assert np.array_equal(fields.floor_gamma(np.zeros((2, 3))), np.full((2, 3), tiny))


def test_generators_are_deterministic():
    a = fields.vahana_wake(20_000)
    b = fields.vahana_wake(20_000)
    assert all(np.array_equal(u, v) for u, v in zip(a, b))

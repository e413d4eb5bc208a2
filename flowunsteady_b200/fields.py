"""Synthetic particle fields for the BASELINE.json configurations.  Deterministic; numpy only.

Every generator returns (X (n,3), Gamma (n,3), sigma (n,)) in float64.  The Julia stack that produces the real
wakes cannot run here, so each generator reproduces the *geometry, particle count and core size* of the reference
example it stands in for; citations give where those numbers come from in /root/reference.
"""
from __future__ import annotations

import numpy as np


def floor_gamma(G: np.ndarray) -> np.ndarray:
    """FLOWUnsteady's add_particle floors every component of Gamma to 5 eps ("or ExaFMM will blow up",
    src/FLOWUnsteady_simulation.jl:464-468); without it a zero-strength particle makes the integrator's 1/|Gamma|^2 a 0/0."""
    tiny = 5 * np.finfo(np.float64).eps
    G = np.array(G, dtype=np.float64, copy=True)
    G[np.abs(G) < tiny] = tiny
    return G


def random_field(n: int, seed: int | None = None):
    """Config 5: x ~ U[0,1)^3, Gamma ~ N(0,1)^3 / n, sigma = 2.125 n^(-1/3) (overlap-preserving), rng seed = n."""
    rng = np.random.default_rng(n if seed is None else seed)
    x = rng.random((n, 3))
    g = rng.standard_normal((n, 3)) / n
    s = np.full(n, 2.125 * n ** (-1.0 / 3.0))
    return x, g, s


def _ring_section(nc: int, a_c: float):
    """Concentric-layer discretisation of a disc of radius a_c: 1 + 3 nc (nc + 1) cells.
    Returns (rho, theta, area) per cell."""
    rl = a_c / (2 * nc + 1)
    rho, th, area = [0.0], [0.0], [np.pi * rl * rl]
    for n in range(1, nc + 1):
        k = np.arange(6 * n)
        rho += [2 * n * rl] * (6 * n)
        th += list(2 * np.pi * k / (6 * n))
        area += [np.pi * ((2 * n + 1) ** 2 - (2 * n - 1) ** 2) * rl * rl / (6 * n)] * (6 * n)
    return np.array(rho), np.array(th), np.array(area)


def vortex_rings(n_total: int = 1_000_000, nrings: int = 2, R: float = 1.0, a: float = 0.1, separation: float = 1.0,
                 Gamma0: float = 1.0, overlap: float = 2.125, disc: float = 2.0):
    """Config 3 (headline): coaxial vortex rings for the leapfrog case, discretised with Nphi cross-sections of
    1 + 3 nc (nc + 1) cells each; nc and Nphi are chosen for near-isotropic spacing h and the field is padded with
    (5 eps)-strength particles on the axis to exactly n_total.  Gaussian vorticity of core size `a` over a disc of
    radius disc*a; sigma = overlap * h (lambda = 2.125 as in examples/rotorhover/rotorhover.jl:155)."""
    a_c = disc * a
    per_ring = n_total / nrings
    # isotropic spacing: (2 pi R / h) * (pi a_c^2 / h^2) = per_ring
    h = (2 * np.pi * R * np.pi * a_c * a_c / per_ring) ** (1.0 / 3.0)
    nc = max(1, int(round((a_c / (h / 2) - 1) / 2)))
    cells = 1 + 3 * nc * (nc + 1)
    nphi = max(8, int(per_ring // cells))
    rho, th, area = _ring_section(nc, a_c)
    w = np.exp(-(rho / a) ** 2) * area
    w *= Gamma0 / w.sum()                       # circulation carried by each cell
    phi = 2 * np.pi * (np.arange(nphi) + 0.5) / nphi
    X, G = [], []
    for r in range(nrings):
        z0 = r * separation
        # cell position in the (radial, axial) plane of each section
        rad = R + rho[None, :] * np.cos(th[None, :])           # (1, cells)
        zz = z0 + rho[None, :] * np.sin(th[None, :])
        cphi, sphi = np.cos(phi)[:, None], np.sin(phi)[:, None]
        x = rad * cphi
        y = rad * sphi
        z = np.broadcast_to(zz, x.shape)
        ds = 2 * np.pi * rad / nphi                            # arc length of the cell's filament segment
        gmag = w[None, :] * ds
        gx = -sphi * gmag
        gy = cphi * gmag
        gz = np.zeros_like(gx)
        X.append(np.stack([x, y, z], -1).reshape(-1, 3))
        G.append(np.stack([gx, gy, gz], -1).reshape(-1, 3))
    X = np.concatenate(X)
    G = np.concatenate(G)
    h_eff = max(2 * a_c / (2 * nc + 1), 2 * np.pi * R / nphi)
    sigma = np.full(X.shape[0], overlap * h_eff)
    npad = n_total - X.shape[0]
    if npad < 0:
        X, G, sigma = X[:n_total], G[:n_total], sigma[:n_total]
    elif npad > 0:
        zp = np.linspace(-separation, (nrings + 1) * separation, npad)
        Xp = np.stack([np.zeros(npad), np.zeros(npad), zp], -1)
        X = np.concatenate([X, Xp])
        # zero-strength padding: floored to 5 eps per component exactly as the reference floors shed particles
        # ("or ExaFMM will blow up", src/FLOWUnsteady_simulation.jl:464-468), so the integrator's 1/|Gamma|^2 is finite
        G = np.concatenate([G, np.full((npad, 3), 5 * np.finfo(np.float64).eps)])
        sigma = np.concatenate([sigma, np.full(npad, overlap * h_eff)])
    return np.ascontiguousarray(X), np.ascontiguousarray(G), np.ascontiguousarray(sigma)


def wing_wake(rows: int = 100, nspan: int = 101, b: float = 2.489, Vinf: float = 49.7, lam: float = 2.0,
              nsteps: int = 200, wakelength: float = 2.75, Gamma0: float = 1.0):
    """Config 1 stand-in (examples/wing): flat wake sheet, `nspan` spanwise filaments x `rows` streamwise rows
    (N = 10,100 / 20,200 / 40,401 in the example: examples/wing/wing.jl:46,58,147), sigma = lambda V dt with
    dt = wakelength*b/Vinf/nsteps (wing.jl:29,37,56-64) = 0.0684 m; streamwise Gamma of elliptic spanwise strength."""
    dt = wakelength * b / Vinf / nsteps
    dx = Vinf * dt
    sigma = lam * dx
    y = np.linspace(-b / 2, b / 2, nspan)
    x = dx * (np.arange(rows) + 1)
    Xg, Yg = np.meshgrid(x, y, indexing="ij")
    # trailing vorticity = -d(Gamma_bound)/dy of an elliptic loading
    eta = np.clip(2 * y / b, -0.999, 0.999)
    dGdy = Gamma0 * eta / np.sqrt(1 - eta * eta) * (2 / b)
    gx = np.broadcast_to(dGdy * (b / (nspan - 1)) * dx, Xg.shape)
    X = np.stack([Xg.ravel(), Yg.ravel(), 0.02 * np.sin(3 * Xg.ravel())], -1)
    G = floor_gamma(np.stack([gx.ravel(), np.zeros(X.shape[0]), np.zeros(X.shape[0])], -1))
    return np.ascontiguousarray(X), np.ascontiguousarray(G), np.full(X.shape[0], sigma)


def rotor_wake(n_total: int = 70_000, blades: int = 2, R: float = 0.12, nfil: int = 41, nsteps_per_rev: int = 36,
               p_per_step: int = 4, lam: float = 2.125, pitch: float = 0.035, Gamma0: float = 0.05):
    """Config 2 stand-in (examples/rotorhover): helical wake of a B=2, R=0.12 m rotor (database/rotors/DJI9443.csv:2,4),
    2n+1 = 41 or 101 trailing filaments per blade (rotorhover.jl:151) and
    sigma = lambda * 2 pi R / (nsteps_per_rev * p_per_step) (rotorhover.jl:155-157)."""
    sigma = lam * 2 * np.pi * R / (nsteps_per_rev * p_per_step)
    per_fil = max(2, n_total // (blades * nfil))
    dpsi = 2 * np.pi / (nsteps_per_rev * p_per_step)
    psi = dpsi * np.arange(per_fil)
    rfil = R * np.linspace(0.15, 1.0, nfil)
    X, G = [], []
    for bl in range(blades):
        ang = psi[None, :] + 2 * np.pi * bl / blades                  # (1, per_fil)
        rr = rfil[:, None] * (1 - 0.22 * (1 - np.exp(-psi[None, :] / 4)))  # wake contraction
        x = rr * np.cos(ang)
        y = rr * np.sin(ang)
        z = -pitch * R * psi[None, :] * np.ones_like(rr)
        # filament tangent * circulation, tip filaments strongest
        gam = Gamma0 * (rfil[:, None] / R) ** 2
        tx, ty, tz = -rr * np.sin(ang) * dpsi, rr * np.cos(ang) * dpsi, -pitch * R * dpsi * np.ones_like(rr)
        X.append(np.stack([x, y, z], -1).reshape(-1, 3))
        G.append(np.stack([gam * tx, gam * ty, gam * tz], -1).reshape(-1, 3))
    X = np.concatenate(X)[:n_total]
    G = np.concatenate(G)[:n_total]
    return np.ascontiguousarray(X), np.ascontiguousarray(G), np.full(X.shape[0], sigma)


def vahana_wake(n_total: int = 5_000_000, span: float = 5.86, nrotors: int = 8, R: float = 0.75, sigma: float = 0.0366,
                _oversample: float | None = None):
    """Config 4 stand-in (examples/vahana): `nrotors` helical rotor wakes trailing two wings plus two planar wing wakes,
    clipped to the wake-treatment sphere of radius 1.25 b (examples/vahana/vahana.jl:356; b = 5.86 m), with the example's
    core size sigma = lambda V dt / p_per_step ~ 0.0366 m (vahana.jl:109; lambda = 2.125, dt = 30/21600 s, p_per_step = 5).
    Each rotor wake takes 1/(nrotors + 2) of the particles; the wing wakes are flat sheets of the same count.  The
    generator oversamples (pilot run) so that exactly n_total particles remain inside the sphere."""
    if _oversample is None:
        pilot = vahana_wake(40_000, span, nrotors, R, sigma, _oversample=1.0)[0].shape[0]
        _oversample = 1.03 * 40_000 / pilot
    per = int(n_total * _oversample) // (nrotors + 2)
    rsph = 1.25 * span
    length = 1.6 * rsph                                        # wakes leave the sphere before they end
    X, G = [], []
    # rotor wakes: helices along -x (cruise), 2 blades x 21 trailing filaments, advance per revolution = 0.6 R
    nfil, blades, adv = 21, 2, 0.6 * R
    for r in range(nrotors):
        wing = r % 2                                           # two wings (tandem): x offset and height differ
        y0 = (r // 2 - (nrotors // 2 - 1) / 2) * (span / (nrotors // 2))
        x0, z0 = (0.0 if wing == 0 else -0.45 * span), (0.0 if wing == 0 else 0.25 * span)
        per_fil = max(2, per // (blades * nfil))
        psi = np.linspace(0.0, 2 * np.pi * length / adv, per_fil)
        dpsi = psi[1] - psi[0]
        rfil = R * np.linspace(0.2, 1.0, nfil)
        for bl in range(blades):
            ang = psi[None, :] + 2 * np.pi * bl / blades + 0.37 * r
            rr = rfil[:, None] * (1 - 0.12 * (1 - np.exp(-psi[None, :] / 6)))
            x = x0 - adv * psi[None, :] / (2 * np.pi) * np.ones_like(rr)
            y = y0 + rr * np.cos(ang)
            z = z0 + rr * np.sin(ang)
            gam = 2.0 * (rfil[:, None] / R) ** 2
            tx = -adv / (2 * np.pi) * dpsi * np.ones_like(rr)
            ty, tz = -rr * np.sin(ang) * dpsi, rr * np.cos(ang) * dpsi
            X.append(np.stack([x, y, z], -1).reshape(-1, 3))
            G.append(np.stack([gam * tx, gam * ty, gam * tz], -1).reshape(-1, 3))
    # wing wakes: flat sheets behind each wing, elliptic loading
    for wing in range(2):
        nspan = 401
        rows = max(2, per // nspan)
        x0, z0 = (0.0 if wing == 0 else -0.45 * span), (0.0 if wing == 0 else 0.25 * span)
        y = np.linspace(-span / 2, span / 2, nspan)
        xs = x0 - length * (np.arange(rows) + 1) / rows
        Xg, Yg = np.meshgrid(xs, y, indexing="ij")
        eta = np.clip(2 * y / span, -0.999, 0.999)
        dG = 8.0 * eta / np.sqrt(1 - eta * eta) * (2 / span) * (span / (nspan - 1)) * (length / rows)
        X.append(np.stack([Xg.ravel(), Yg.ravel(), z0 + 0.05 * np.sin(2 * Xg.ravel())], -1))
        G.append(np.stack([np.broadcast_to(dG, Xg.shape).ravel(), np.zeros(Xg.size), np.zeros(Xg.size)], -1))
    X = np.concatenate(X)
    G = np.concatenate(G)
    keep = np.linalg.norm(X - np.array([-0.2 * span, 0.0, 0.1 * span]), axis=1) < rsph   # remove_particles_sphere
    X, G = X[keep], G[keep]
    if X.shape[0] > n_total:                                   # thin uniformly down to n_total
        sel = np.linspace(0, X.shape[0] - 1, n_total).astype(np.int64)
        X, G = X[sel], G[sel]
    return np.ascontiguousarray(X), np.ascontiguousarray(floor_gamma(G)), np.full(X.shape[0], sigma)


def ring_impulse(X: np.ndarray, G: np.ndarray) -> np.ndarray:
    """Linear impulse 0.5 sum x_p x Gamma_p — conserved by the inviscid dynamics (used as a size-independent check)."""
    return 0.5 * np.cross(X, G).sum(axis=0)

"""Shared builders for the parity tests (inputs come from the oracle's seeded generators)."""
import numpy as np

from oracle import oracle as orc

LIQUID_CELL = 0.36165          # rho*sigma^3 ≈ 0.844
LONG_CUT = (1.1963, -0.003723224030513348)  # 3.5 sigma potentials.json entry


def f8(x):
    s = f"{x:.8f}"
    return "0.00000000" if s == "-0.00000000" else s


def fmt3(v):
    return [f8(float(c)) for c in v]


def liquid(n_side=8, seed=3, jitter=0.03, temperature=120.0):
    st = orc.argon_lattice(n_side, LIQUID_CELL, temperature, seed)
    rng = np.random.default_rng(seed)
    st.pos += rng.uniform(-jitter, jitter, st.pos.shape)
    orc.apply_boundary_conditions(st)
    return st


def gas(n_side=10, seed=42, temperature=273.15):
    return orc.argon_lattice(n_side, orc.GAS_CELL, temperature, seed)


def dense_gas(n=3000, box=(9.0, 7.5, 8.1), seed=5, temperature=273.15):
    """Random positions with a minimum separation so no pair sits deep inside the repulsive core."""
    rng = np.random.default_rng(seed)
    box = np.array(box)
    pos = []
    grid = {}
    rmin = 0.30
    cell = rmin
    nc = np.maximum((box / cell).astype(int), 1)
    while len(pos) < n:
        p = rng.uniform(0, 1, 3) * box
        c = tuple((p / box * nc).astype(int) % nc)
        ok = True
        for dx in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dz in (-1, 0, 1):
                    key = ((c[0] + dx) % nc[0], (c[1] + dy) % nc[1], (c[2] + dz) % nc[2])
                    for q in grid.get(key, ()):
                        d = p - q
                        d -= box * np.round(d / box)
                        if d @ d < rmin * rmin:
                            ok = False
        if ok:
            grid.setdefault(c, []).append(p)
            pos.append(p)
    pos = np.array(pos)
    vel = orc.init_velocities(n, temperature, orc.ARGON_MASS, seed)
    return orc.State(pos, vel, orc.ARGON_MASS, box)


def to_gpu_state(md, st):
    """oracle State → moldyn_b200 State (same arrays, copied)."""
    g = md.State(st.pos, st.vel, st.mass, st.box)
    g.force[:] = st.force
    g.potential[:] = st.pot
    g.temp[:] = st.vir
    return g


def lj_pair(md, r_cut=None, u_cut=None):
    """(oracle potential, product potential) with identical parameters."""
    if r_cut is None:
        o = orc.LennardJones()
        p = md.Potential.new_lennard_jones(0.3418, 1.712)
    else:
        o = orc.LennardJones(r_cut=r_cut, u_cut=u_cut)
        p = md.Potential(0.3418, 1.712, r_cut, u_cut)
    return o, p


def force_scale(lj, st):
    """Per-atom Σ_j |f_ij| (the scale relative errors are measured against: on a lattice ΣF≈0)."""
    off, nbr = orc.neighbour_sets(st.pos, st.box, lj.r_cut)
    scale = np.zeros(st.n)
    for i in range(st.n):
        js = nbr[off[i]:off[i + 1]]
        if len(js) == 0:
            continue
        d = st.pos[js] - st.pos[i]
        for k in range(3):
            L = st.box[k]
            d[:, k] = np.where(d[:, k] < -L / 2.0, d[:, k] + L, np.where(d[:, k] > L / 2.0, d[:, k] - L, d[:, k]))
        r = np.sqrt((d * d).sum(axis=1))
        s6 = (lj.sigma / r) ** 6
        scale[i] = np.abs(24.0 * lj.eps / r * (s6 - 2.0 * s6 * s6)).sum()
    return scale

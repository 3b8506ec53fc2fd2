"""Several particle types: the oracle's restatement of what the reference does (potential.rs:158-216, integrator.rs:14-59
with State.particles of length T > 1) pinned against an independent pure-Python restatement written here from the same
reference lines, and against the single-type oracle where the two must coincide."""
import ctypes
import math

import numpy as np

from oracle import oracle as orc

K_B = 1.380648528
_libm = ctypes.CDLL("libm.so.6")
_libm.cbrt.restype = ctypes.c_double
_libm.cbrt.argtypes = [ctypes.c_double]
_cbrt = _libm.cbrt


def mixture(n_side=4, cell=0.45, counts=(20, 30, 14), masses=(66.335, 20.18, 131.29), temperature=150.0, seed=3, jitter=0.05):
    st = orc.argon_lattice(n_side, cell, temperature, seed)
    rng = np.random.default_rng(seed)
    perm = rng.permutation(st.n)            # types interleaved in space
    pos = st.pos[perm] + rng.uniform(-jitter, jitter, st.pos.shape)
    m = orc.MultiState(pos, st.vel[perm], counts, masses, st.box)
    assert m.n == sum(counts)
    orc.lib().orc_apply_boundary_conditions(m.n, orc._p(m.pos), orc._p(m.box))
    return m


def table3():
    tab = orc.PotentialTable(3)
    tab.set_potential(0, 1, orc.LennardJones(0.31, 1.1))
    tab.set_potential(1, 1, orc.LennardJones(0.28, 0.6, r_cut=0.8))
    tab.set_potential(2, 0, orc.LennardJones(0.37, 2.0))
    # (1, 2) and (2, 2) are not set: PotentialsDatabase::get_potential falls back to the default (potential.rs:147-155)
    return tab


# ---- pure-Python restatement (floats are IEEE doubles, no contraction) ---------------------------------------------
def py_lj(p, r):  # potential.rs:57-70
    if r > p.r_cut:
        return 0.0, 0.0
    s = p.sigma / r
    x2 = s * s
    x4 = x2 * x2
    s6 = x2 * x4
    s12 = s6 * s6
    return 4.0 * p.eps * (s12 - s6) - p.u_cut, (24.0 * p.eps / r) * (s6 - 2.0 * s12)


def py_update_force(tab, st, symmetric):  # potential.rs:158-216
    T, bb = st.T, [float(x) for x in st.box]
    F = [[0.0, 0.0, 0.0] for _ in range(st.n)]
    U = [0.0] * st.n
    W = [0.0] * st.n
    P = [[float(c) for c in row] for row in st.pos]
    for t1 in range(T):
        for t2 in (range(T) if symmetric else range(t1, T)):
            p = tab.get_potential(t1, t2)
            for i in range(st.start[t1], st.start[t1 + 1]):
                for j in range(st.start[t2], st.start[t2 + 1]):
                    if i == j:
                        continue
                    r = [P[j][d] - P[i][d] for d in range(3)]
                    for d in range(3):
                        if r[d] < -bb[d] / 2.0:
                            r[d] += bb[d]
                        elif r[d] > bb[d] / 2.0:
                            r[d] -= bb[d]
                    r_abs = math.sqrt((r[0] * r[0] + r[1] * r[1]) + r[2] * r[2])
                    if r_abs > p.r_cut:
                        continue
                    u, f = py_lj(p, r_abs)
                    fv = [r[d] / r_abs * f for d in range(3)]
                    t = fv[0] * r[0] + fv[1] * r[1] + fv[2] * r[2]
                    for d in range(3):
                        F[i][d] += fv[d]
                    U[i] += u
                    W[i] += t
    return np.array(F), np.array(U), np.array(W)


def py_type_macro(st, t, vel, vir, bb):  # macro_parameters/{mod,energy,temperature,pressure}.rs, one type
    a, b, m = int(st.start[t]), int(st.start[t + 1]), float(st.masses[t])
    s = [0.0, 0.0, 0.0, 0.0]
    for i in range(a, b):
        for d in range(3):
            s[d] += vel[i][d] * m
        s[3] += 1.0 * m
    mv = [s[d] / s[3] for d in range(3)]
    th = 0.0
    r1 = r2 = 0.0
    for i in range(a, b):
        dv = [vel[i][d] - mv[d] for d in range(3)]
        th += m * ((dv[0] * dv[0] + dv[1] * dv[1]) + dv[2] * dv[2]) / 2.0
        r1 += m * dv[0] * dv[0]
        r1 += m * dv[1] * dv[1]
        r1 += m * dv[2] * dv[2]
        r2 -= vir[i]
    temperature = (2.0 * th) / (3.0 * float(b - a) * K_B) * 100.0
    pressure = (r1 + r2 * 0.5) / (bb[0] * bb[1] * bb[2]) / 3.0
    return temperature, pressure


def py_step(tab, st, dt, th, ba, symmetric):
    """integrator.rs:14-59 with Berendsen thermostat (tau, T0) and barostat (beta, tau, P0); returns lambda, myu."""
    T = st.T
    vel = [[float(c) for c in row] for row in st.vel]
    pos = [[float(c) for c in row] for row in st.pos]
    frc = [[float(c) for c in row] for row in st.force]
    vir = [float(x) for x in st.vir]
    bb = [float(x) for x in st.box]
    lam = myu = None
    if ba:
        for t in range(T):
            _, pr = py_type_macro(st, t, vel, vir, bb)
            myu = _cbrt(1.0 + dt * ba[0] / ba[1] * (pr - ba[2]))  # f64::cbrt = libm's cbrt
    if th:
        for t in range(T):
            te, _ = py_type_macro(st, t, vel, vir, bb)
            lam = math.sqrt(1.0 + dt / th[0] * (th[1] / te - 1.0))
    for t in range(T):
        c = dt / (2.0 * float(st.masses[t]))
        for i in range(st.start[t], st.start[t + 1]):
            for d in range(3):
                vel[i][d] = vel[i][d] + frc[i][d] * c
    if th:
        for i in range(st.n):
            for d in range(3):
                vel[i][d] *= lam
    for i in range(st.n):
        for d in range(3):
            pos[i][d] += vel[i][d] * dt
            if pos[i][d] < 0.0:
                pos[i][d] += bb[d]
            elif pos[i][d] >= bb[d]:
                pos[i][d] -= bb[d]
    st.pos[:] = np.array(pos)
    F, U, W = py_update_force(tab, st, symmetric)
    for t in range(T):
        c = dt / (2.0 * float(st.masses[t]))
        for i in range(st.start[t], st.start[t + 1]):
            for d in range(3):
                vel[i][d] += float(F[i][d]) * c
    if ba:
        for t in range(T):
            bb = [x * myu for x in bb]
            for i in range(st.start[t], st.start[t + 1]):
                for d in range(3):
                    pos[i][d] *= myu
    st.pos[:] = np.array(pos)
    st.vel[:] = np.array(vel)
    st.force[:] = F
    st.pot[:] = U
    st.vir[:] = W
    st.box[:] = bb
    return lam, myu


# ---- tests ---------------------------------------------------------------------------------------------------------
def test_two_atoms_two_types_reference_is_one_sided():
    pos = np.array([[1.0, 1.0, 1.0], [1.4, 1.0, 1.0]])
    st = orc.MultiState(pos, np.zeros((2, 3)), [1, 1], [66.335, 30.0], [10, 10, 10])
    tab = orc.PotentialTable(2)
    u, f = orc.LennardJones().get_potential_and_force(1.4 - 1.0)
    orc.update_force_multi(tab, st)
    # potential.rs:168-176: type 0 accumulates its type-1 partner, type 1 never sees type 0
    assert np.array_equal(st.force, [[f, 0.0, 0.0], [0.0, 0.0, 0.0]]) and np.array_equal(st.pot, [u, 0.0])
    orc.update_force_multi(tab, st, symmetric=True)
    assert np.array_equal(st.force, [[f, 0.0, 0.0], [-f, 0.0, 0.0]]) and np.array_equal(st.pot, [u, u])


def test_one_type_is_the_single_type_oracle():
    s1 = orc.argon_lattice(5, 0.5, 120.0, 3)
    m1 = orc.MultiState(s1.pos, s1.vel, [s1.n], [orc.ARGON_MASS], s1.box)
    lj = orc.LennardJones()
    orc.update_force(lj, s1)
    orc.update_force_multi(orc.PotentialTable(1), m1)
    assert np.array_equal(s1.force, m1.force) and np.array_equal(s1.pot, m1.pot) and np.array_equal(s1.vir, m1.vir)
    th = [orc.Thermostat(1, 10.0, 300.0) for _ in range(2)]
    ba = [orc.Barostat(1.0, 5.0, 1.0) for _ in range(2)]
    orc.step(lj, s1, 0.002, th[0], ba[0], n_steps=5)
    orc.step_multi(orc.PotentialTable(1), m1, 0.002, th[1], ba[1], n_steps=5)
    assert np.array_equal(s1.pos, m1.pos) and np.array_equal(s1.vel, m1.vel) and np.array_equal(s1.box, m1.box)
    assert th[0].lambda_ == th[1].lambda_ and ba[0].myu == ba[1].myu
    # the same atoms split into two types with the same mass and the default potentials: symmetric forces are the
    # single-type ones (same ascending-j order)
    m2 = orc.MultiState(s1.pos, s1.vel, [60, s1.n - 60], [orc.ARGON_MASS] * 2, s1.box)
    orc.update_force(lj, s1)
    orc.update_force_multi(orc.PotentialTable(2), m2, symmetric=True)
    assert np.array_equal(s1.force, m2.force) and np.array_equal(s1.pot, m2.pot)


def test_update_force_three_types_against_python():
    tab = table3()
    for symmetric in (False, True):
        st = mixture()
        orc.update_force_multi(tab, st, symmetric=symmetric)
        F, U, W = py_update_force(tab, st, symmetric)
        assert np.array_equal(st.force, F) and np.array_equal(st.pot, U) and np.array_equal(st.vir, W)
        assert np.abs(F).max() > 1.0
        if not symmetric:  # the last type only feels itself
            sub = orc.State(st.pos[st.start[2]:], st.vel[st.start[2]:], st.masses[2], st.box)
            orc.update_force(tab.get_potential(2, 2), sub)
            assert np.array_equal(sub.force, st.force[st.start[2]:])


def test_step_three_types_against_python():
    tab = table3()
    for symmetric in (False, True):
        a = mixture(counts=(12, 9, 6), n_side=3)
        orc.update_force_multi(tab, a, symmetric=symmetric)
        b = a.copy()
        th, ba = orc.Thermostat(1, 0.5, 200.0), orc.Barostat(1.0e-3, 2.0, 1.0)
        for _ in range(4):
            box0 = a.box.copy()
            orc.step_multi(tab, a, 0.002, th, ba, symmetric=symmetric)
            lam, myu = py_step(tab, b, 0.002, (0.5, 200.0), (1.0e-3, 2.0, 1.0), symmetric)
            assert th.lambda_ == lam and ba.myu == myu
            assert np.array_equal(a.pos, b.pos) and np.array_equal(a.vel, b.vel) and np.array_equal(a.box, b.box)
            assert np.array_equal(a.force, b.force) and np.array_equal(a.vir, b.vir)
            # barostat.update runs once per type: the box is scaled T times (integrator.rs:54-58)
            assert np.array_equal(a.box, ((box0 * myu) * myu) * myu) and myu != 1.0
        # the coefficients are those of the LAST type (integrator.rs:18-27)
        c = a.copy()
        t_last = orc.macro_type(c, 2)["temperature"]
        th2 = orc.Thermostat(1, 0.5, 200.0)
        orc.step_multi(tab, c, 0.002, th2, None, symmetric=symmetric)
        assert th2.lambda_ == math.sqrt(1.0 + 0.002 / 0.5 * (200.0 / t_last - 1.0))

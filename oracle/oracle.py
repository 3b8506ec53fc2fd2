"""ctypes front-end of the CPU oracle (oracle/moldyn_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / ``--impl reference`` legs of bench.py — never by moldyn_b200/.
Each method cites the reference lines its C counterpart restates.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libmoldyn_oracle.so")

K_B = 1.380648528  # core/src/lib.rs:15


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "moldyn_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libmoldyn_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _LJ(C.Structure):
    _fields_ = [("sigma", C.c_double), ("eps", C.c_double), ("r_cut", C.c_double), ("u_cut", C.c_double)]


class _Thermostat(C.Structure):
    _fields_ = [("kind", C.c_int), ("tau", C.c_double), ("target", C.c_double),
                ("lambda_", C.c_double), ("psi", C.c_double)]


class _Barostat(C.Structure):
    _fields_ = [("kind", C.c_int), ("beta", C.c_double), ("tau", C.c_double),
                ("target", C.c_double), ("myu", C.c_double)]


class _State(C.Structure):
    _fields_ = [("n", C.c_int64), ("mass", C.c_double),
                ("pos", C.c_void_p), ("vel", C.c_void_p), ("force", C.c_void_p),
                ("pot", C.c_void_p), ("vir", C.c_void_p), ("bb", C.c_double * 3)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_kinetic_energy.restype = C.c_double
        _lib.orc_thermal_energy.restype = C.c_double
        _lib.orc_potential_energy.restype = C.c_double
        _lib.orc_temperature.restype = C.c_double
        _lib.orc_pressure.restype = C.c_double
        _lib.orc_init_positions.restype = C.c_int64
        _lib.orc_num_threads.restype = C.c_int
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class LennardJones:
    """Potential::LennardJones (solver/src/solver/potential.rs:12-87)."""

    def __init__(self, sigma=0.3418, eps=1.712, r_cut=None, u_cut=None):
        self._s = _LJ()
        lib().orc_lj_new(C.c_double(sigma), C.c_double(eps), C.byref(self._s))  # potential.rs:27-55
        if r_cut is not None:  # hand-written potentials.json entry (potential.rs:125-139)
            self._s.r_cut = r_cut
            self._s.u_cut = u_cut if u_cut is not None else self.unshifted_potential(r_cut)
        elif u_cut is not None:
            self._s.u_cut = u_cut

    sigma = property(lambda s: s._s.sigma)
    eps = property(lambda s: s._s.eps)
    r_cut = property(lambda s: s._s.r_cut)
    u_cut = property(lambda s: s._s.u_cut)

    def unshifted_potential(self, r):
        tmp = _LJ(self._s.sigma, self._s.eps, float("inf"), 0.0)
        u, f = C.c_double(), C.c_double()
        lib().orc_lj_potential_and_force(C.byref(tmp), C.c_double(r), C.byref(u), C.byref(f))
        return u.value

    def get_potential_and_force(self, r):  # potential.rs:57-70
        u, f = C.c_double(), C.c_double()
        lib().orc_lj_potential_and_force(C.byref(self._s), C.c_double(r), C.byref(u), C.byref(f))
        return u.value, f.value

    def get_radius_cut(self):  # potential.rs:77-86
        return self._s.r_cut


class State:
    """Single-type restatement of core::State (core/src/particle.rs:6-32), SoA-of-Vector3."""

    def __init__(self, pos, vel, mass, box):
        self.pos = _f64(pos).reshape(-1, 3).copy()
        self.vel = _f64(vel).reshape(-1, 3).copy()
        self.n = self.pos.shape[0]
        assert self.vel.shape == self.pos.shape
        self.mass = float(mass)
        self.box = _f64(box).reshape(3).copy()
        self.force = np.zeros_like(self.pos)
        self.pot = np.zeros(self.n)
        self.vir = np.zeros(self.n)  # Particle.temp = Σ F_ij·r_ij (particle.rs:17-18)

    def copy(self):
        s = State(self.pos, self.vel, self.mass, self.box)
        s.force[:] = self.force
        s.pot[:] = self.pot
        s.vir[:] = self.vir
        return s

    def _c(self):
        st = _State()
        st.n = self.n
        st.mass = self.mass
        st.pos, st.vel, st.force = _p(self.pos), _p(self.vel), _p(self.force)
        st.pot, st.vir = _p(self.pot), _p(self.vir)
        st.bb[:] = list(self.box)
        return st


class Thermostat:
    """Thermostat::{Berendsen,NoseHoover} (solver/src/initializer/thermostat.rs:4-73)."""
    BERENDSEN, NOSE_HOOVER = 1, 2

    def __init__(self, kind, tau, target):
        self._s = _Thermostat(kind, tau, target, 0.0, 0.0)

    lambda_ = property(lambda s: s._s.lambda_)
    psi = property(lambda s: s._s.psi)


class Barostat:
    """Barostat::Berendsen (solver/src/initializer/barostat.rs:4-57)."""
    BERENDSEN = 1

    def __init__(self, beta, tau, target):
        self._s = _Barostat(1, beta, tau, target, 0.0)

    myu = property(lambda s: s._s.myu)


def update_force(lj: LennardJones, st: State, mode="n2", rows=None):
    """update_force (potential.rs:158-216). mode 'n2' = the reference's scan, 'cells' = Θ(N) with
    bit-identical ascending-j sums. rows=(i0,i1) evaluates only those rows (sampling at 1M atoms)."""
    bb = _f64(st.box)
    if rows is not None:
        lib().orc_update_force_rows(C.byref(lj._s), C.c_int64(st.n), _p(st.pos), _p(bb),
                                    C.c_int64(rows[0]), C.c_int64(rows[1]),
                                    _p(st.force), _p(st.pot), _p(st.vir))
    elif mode == "n2":
        lib().orc_update_force(C.byref(lj._s), C.c_int64(st.n), _p(st.pos), _p(bb),
                               _p(st.force), _p(st.pot), _p(st.vir))
    else:
        lib().orc_update_force_cells(C.byref(lj._s), C.c_int64(st.n), _p(st.pos), _p(bb),
                                     _p(st.force), _p(st.pot), _p(st.vir))


def step(lj, st: State, dt, thermostat: Thermostat | None = None, barostat: Barostat | None = None,
         mode="n2", n_steps=1):
    """Integrator::VerletMethod.calculate (solver/src/solver/integrator.rs:14-59), n_steps times."""
    cs = st._c()
    th = C.byref(thermostat._s) if thermostat is not None else None
    ba = C.byref(barostat._s) if barostat is not None else None
    for _ in range(n_steps):
        lib().orc_step(C.byref(lj._s), C.byref(cs), C.c_double(dt), th, ba, C.c_int(0 if mode == "n2" else 1))
    st.box[:] = list(cs.bb)


def apply_boundary_conditions(st: State):  # core/src/particle.rs:120-142
    lib().orc_apply_boundary_conditions(C.c_int64(st.n), _p(st.pos), _p(_f64(st.box)))


def center_of_mass_velocity(st: State):  # macro_parameters/mod.rs:12-25
    out = np.zeros(3)
    lib().orc_center_of_mass_velocity(C.c_int64(st.n), _p(st.vel), C.c_double(st.mass), _p(out))
    return out


def momentum(st: State):  # macro_parameters/mod.rs:28-34
    out = np.zeros(3)
    lib().orc_momentum(C.c_int64(st.n), _p(st.vel), C.c_double(st.mass), _p(out))
    return out


def kinetic_energy(st: State):  # energy.rs:14-22
    return lib().orc_kinetic_energy(C.c_int64(st.n), _p(st.vel), C.c_double(st.mass))


def thermal_energy(st: State, vcom):  # energy.rs:25-37
    return lib().orc_thermal_energy(C.c_int64(st.n), _p(st.vel), C.c_double(st.mass), _p(_f64(vcom)))


def potential_energy(st: State):  # energy.rs:40-49
    return lib().orc_potential_energy(C.c_int64(st.n), _p(st.pot))


def temperature(thermal, n):  # temperature.rs:4-7
    return lib().orc_temperature(C.c_double(thermal), C.c_int64(n))


def pressure(st: State, vcom):  # pressure.rs:5-20
    return lib().orc_pressure(C.c_int64(st.n), _p(st.vel), _p(st.vir), C.c_double(st.mass),
                              _p(_f64(st.box)), _p(_f64(vcom)))


def macro(st: State):
    """All macro parameters of the state, as solve_macro computes them (cli/src/commands.rs:237-264)."""
    vc = center_of_mass_velocity(st)
    th = thermal_energy(st, vc)
    return {"vcom": vc, "kinetic": kinetic_energy(st), "thermal": th, "potential": potential_energy(st),
            "temperature": temperature(th, st.n), "pressure": pressure(st, vc)}


def init_positions(kind, size, cell, start=(0.0, 0.0, 0.0)):
    """UnitCell::{U,FCC}.initialize_particles_position (initializer/position.rs:24-104)."""
    k = 0 if kind in ("u", "U", 0) else 1
    n = size[0] * size[1] * size[2] * (4 if k else 1)
    pos = np.zeros((n, 3))
    got = lib().orc_init_positions(C.c_int(k), C.c_int64(size[0]), C.c_int64(size[1]), C.c_int64(size[2]),
                                   C.c_double(cell), _p(_f64(start)), _p(pos))
    assert got == n
    return pos


def init_velocities(n, temperature_kelvin, mass, seed=42):
    """initialize_velocities_maxwell_boltzmann (initializer/velocity.rs:6-29), seeded."""
    vel = np.zeros((n, 3))
    lib().orc_init_velocities(C.c_int64(n), C.c_double(temperature_kelvin), C.c_double(mass),
                              C.c_uint64(seed), _p(vel))
    return vel


def random_positions(n, box, seed=42):
    pos = np.zeros((n, 3))
    lib().orc_random_positions(C.c_int64(n), _p(_f64(box)), C.c_uint64(seed), _p(pos))
    return pos


def neighbour_sets(pos, box, r_list):
    """CSR (offsets, partners) of {j != i : reference min-image |r_ij| <= r_list} (potential.rs:181-204)."""
    pos = _f64(pos).reshape(-1, 3)
    n = pos.shape[0]
    bb = _f64(box)
    counts = np.zeros(n, dtype=np.int64)
    lib().orc_neighbour_sets(C.c_int64(n), _p(pos), _p(bb), C.c_double(r_list), _p(counts), None, None)
    offsets = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=offsets[1:])
    nbr = np.zeros(max(int(offsets[-1]), 1), dtype=np.int64)
    lib().orc_neighbour_sets(C.c_int64(n), _p(pos), _p(bb), C.c_double(r_list), _p(counts), _p(offsets), _p(nbr))
    return offsets, nbr[: int(offsets[-1])]


# --- several particle types (State.particles indexed by type id, flattened type by type) ----------------------------

class _StateMulti(C.Structure):
    _fields_ = [("T", C.c_int), ("start", C.c_void_p), ("mass", C.c_void_p),
                ("pos", C.c_void_p), ("vel", C.c_void_p), ("force", C.c_void_p),
                ("pot", C.c_void_p), ("vir", C.c_void_p), ("bb", C.c_double * 3)]


class MultiState:
    """core::State with several particle types (core/src/particle.rs:24-32): `counts[t]` atoms of type t, stored type by
    type; `masses[t]` is Particle.mass of the type."""

    def __init__(self, pos, vel, counts, masses, box):
        self.pos = _f64(pos).reshape(-1, 3).copy()
        self.vel = _f64(vel).reshape(-1, 3).copy()
        self.n = self.pos.shape[0]
        self.counts = np.asarray(counts, dtype=np.int64).copy()
        assert self.counts.sum() == self.n and (self.counts > 0).all()
        self.start = np.zeros(len(self.counts) + 1, dtype=np.int64)
        np.cumsum(self.counts, out=self.start[1:])
        self.masses = _f64(masses).copy()
        assert self.masses.shape == self.counts.shape
        self.box = _f64(box).reshape(3).copy()
        self.force = np.zeros_like(self.pos)
        self.pot = np.zeros(self.n)
        self.vir = np.zeros(self.n)

    T = property(lambda s: len(s.counts))

    def types(self):
        """type id of every atom (uint16 like Particle.id)"""
        return np.repeat(np.arange(self.T, dtype=np.uint16), self.counts)

    def copy(self):
        s = MultiState(self.pos, self.vel, self.counts, self.masses, self.box)
        s.force[:] = self.force
        s.pot[:] = self.pot
        s.vir[:] = self.vir
        return s

    def _c(self):
        st = _StateMulti()
        st.T = self.T
        st.start, st.mass = _p(self.start), _p(self.masses)
        st.pos, st.vel, st.force = _p(self.pos), _p(self.vel), _p(self.force)
        st.pot, st.vir = _p(self.pot), _p(self.vir)
        st.bb[:] = list(self.box)
        return st


class PotentialTable:
    """PotentialsDatabase (potential.rs:89-155): entries keyed (min id, max id), everything else the default potential."""

    def __init__(self, n_types, default=None):
        self.T = n_types
        self.default = default or LennardJones()
        self.entries = {}

    def set_potential(self, id0, id1, lj):  # potential.rs:141-144
        self.entries[(min(id0, id1), max(id0, id1))] = lj

    def get_potential(self, id0, id1):  # potential.rs:147-155
        return self.entries.get((min(id0, id1), max(id0, id1)), self.default)

    def _c(self):
        arr = (_LJ * (self.T * self.T))()
        for a in range(self.T):
            for b in range(self.T):
                lj = self.get_potential(a, b)._s
                arr[a * self.T + b] = _LJ(lj.sigma, lj.eps, lj.r_cut, lj.u_cut)
        return arr


def update_force_multi(table: PotentialTable, st: MultiState, symmetric=False):
    """update_force with several types (potential.rs:158-216).  symmetric=False is the reference: atoms of type t1 only
    accumulate partners of types t2 >= t1."""
    lib().orc_update_force_multi(C.c_int(st.T), _p(st.start), table._c(), _p(st.pos), _p(_f64(st.box)),
                                 C.c_int(1 if symmetric else 0), _p(st.force), _p(st.pot), _p(st.vir))


def step_multi(table: PotentialTable, st: MultiState, dt, thermostat: Thermostat | None = None,
               barostat: Barostat | None = None, symmetric=False, n_steps=1):
    """Integrator::VerletMethod.calculate with several types (integrator.rs:14-59): per-type calculate_myu /
    calculate_lambda (the last type's coefficient is applied to all), per-type masses, box scaled once per type."""
    cs = st._c()
    tab = table._c()
    th = C.byref(thermostat._s) if thermostat is not None else None
    ba = C.byref(barostat._s) if barostat is not None else None
    for _ in range(n_steps):
        lib().orc_step_multi(tab, C.byref(cs), C.c_double(dt), th, ba, C.c_int(1 if symmetric else 0))
    st.box[:] = list(cs.bb)


def macro_type(st: MultiState, t):
    """The reference's per-type macro parameters (macro_parameters/*.rs all take particle_type_id)."""
    out = np.zeros(8)
    lib().orc_macro_type(C.byref(st._c()), C.c_int(t), _p(out))
    return {"kinetic": out[0], "thermal": out[1], "potential": out[2], "temperature": out[3], "pressure": out[4],
            "vcom": out[5:8].copy()}


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(t):
    lib().orc_set_num_threads(C.c_int(t))


# --- the argon systems of BASELINE.json / SURVEY §8d ---------------------------------------------
ARGON_MASS = 66.335
ARGON_RADIUS = 0.071
GAS_CELL = 3.338339


def argon_lattice(size, cell=GAS_CELL, temperature_kelvin=273.15, seed=42, kind="u"):
    """`moldyn-cli initialize -t u -s size -l cell -T T` (cli/src/commands.rs:43-80) with a seeded RNG."""
    if isinstance(size, int):
        size = (size, size, size)
    pos = init_positions(kind, size, cell)
    vel = init_velocities(pos.shape[0], temperature_kelvin, ARGON_MASS, seed)
    box = np.array([cell * size[0], cell * size[1], cell * size[2]])
    return State(pos, vel, ARGON_MASS, box)

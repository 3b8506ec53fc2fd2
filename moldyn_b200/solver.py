"""Host-side mirror of the reference's solver surface for the `solve` hot path, forwarding to the C ABI.

Names, argument meaning and error behaviour follow the reference (paths relative to AndrewChe7/moldyn):
    Potential / PotentialsDatabase / update_force      solver/src/solver/potential.rs
    Integrator                                         solver/src/solver/integrator.rs
    Thermostat / Barostat                              solver/src/initializer/{thermostat,barostat}.rs
    State                                              core/src/particle.rs
    get_* macro parameters                             solver/src/macro_parameters/*.rs
All arithmetic runs in the CUDA library (include/moldyn_b200.h); nothing here computes physics on the CPU
and there is no fallback when the library or a GPU is missing.
"""
from __future__ import annotations

import ctypes as C
import json
import os

import numpy as np

from . import _ffi
from ._ffi import MdError  # noqa: F401  (re-export)

K_B = 1.380648528  # core/src/lib.rs:15


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------------------------------
class Potential:
    """Potential::LennardJones { sigma, eps, r_cut, u_cut } (potential.rs:12-23)."""

    def __init__(self, sigma, eps, r_cut, u_cut):
        self.sigma, self.eps, self.r_cut, self.u_cut = float(sigma), float(eps), float(r_cut), float(u_cut)

    @staticmethod
    def new_lennard_jones(sigma, eps):  # potential.rs:27-55
        rc, uc = C.c_double(), C.c_double()
        _ffi.check(None, _ffi.lib().md_lj_new(sigma, eps, C.byref(rc), C.byref(uc)))
        return Potential(sigma, eps, rc.value, uc.value)

    def get_potential_and_force(self, r):  # potential.rs:57-70
        u, f = C.c_double(), C.c_double()
        _ffi.check(None, _ffi.lib().md_lj_potential_and_force(self.sigma, self.eps, self.r_cut, self.u_cut, r,
                                                             C.byref(u), C.byref(f)))
        return u.value, f.value

    def get_radius_cut(self):  # potential.rs:77-86
        return self.r_cut

    def to_json(self):  # serde externally-tagged enum (potential.rs:11-18)
        return {"LennardJones": {"sigma": self.sigma, "eps": self.eps, "r_cut": self.r_cut, "u_cut": self.u_cut}}

    @staticmethod
    def from_json(obj):
        if "LennardJones" not in obj:
            raise MdError(4, "Potential::Custom is todo!() in the reference")
        d = obj["LennardJones"]
        return Potential(d["sigma"], d["eps"], d["r_cut"], d["u_cut"])


class PotentialsDatabase:
    """PotentialsDatabase (potential.rs:89-155): (min id, max id) → Potential, default = argon LJ."""

    def __init__(self):  # PotentialsDatabase::new  potential.rs:95-101
        self.potentials = {}
        self.default_potential = Potential.new_lennard_jones(0.3418, 1.712)

    def set_potential(self, id0, id1, potential):  # potential.rs:141-144
        self.potentials[(min(id0, id1), max(id0, id1))] = potential

    def get_potential(self, id0, id1):  # potential.rs:147-154
        return self.potentials.get((min(id0, id1), max(id0, id1)), self.default_potential)

    def save_potentials_to_file(self, path):  # potential.rs:104-122
        os.makedirs(path, exist_ok=True)
        with open(os.path.join(path, "potentials.json"), "w") as f:
            json.dump({f"{k[0]},{k[1]}": v.to_json() for k, v in self.potentials.items()}, f, indent=2)

    def load_potentials_from_file(self, path):  # potential.rs:125-139
        with open(os.path.join(path, "potentials.json")) as f:
            data = json.load(f)
        for key, val in data.items():
            a, b = (int(x) for x in key.split(","))
            self.potentials[(a, b)] = Potential.from_json(val)


class Thermostat:
    """Thermostat enum (thermostat.rs:4-22). lambda (and psi) are stored back after every step."""

    def __init__(self, kind, tau):
        self.kind, self.tau, self.lambda_, self.psi = kind, float(tau), 0.0, 0.0

    @staticmethod
    def Berendsen(tau, lambda_=0.0):
        return Thermostat(_ffi.THERMOSTAT_BERENDSEN, tau)

    @staticmethod
    def NoseHoover(tau, psi=0.0, lambda_=0.0):
        t = Thermostat(_ffi.THERMOSTAT_NOSE_HOOVER, tau)
        t.psi = psi
        return t

    def _c(self, target):
        return _ffi.ThermostatC(self.kind, 0, self.tau, float(target), self.lambda_, self.psi)


class Barostat:
    """Barostat enum (barostat.rs:4-19). myu is stored back after every step."""

    def __init__(self, kind, beta, tau):
        self.kind, self.beta, self.tau, self.myu = kind, float(beta), float(tau), 0.0

    @staticmethod
    def Berendsen(beta, tau, myu=0.0):
        return Barostat(_ffi.BAROSTAT_BERENDSEN, beta, tau)

    def _c(self, target):
        return _ffi.BarostatC(self.kind, 0, self.beta, self.tau, float(target), self.myu)


class State:
    """core::State for ONE particle type (particle.rs:25-32): `particles[0]` as flat f64 arrays plus
    `boundary_box`.  States with several types: MultiState."""

    def __init__(self, position, velocity, mass, boundary_box, radius=0.1, particle_id=0):
        self.position = _f64(position).reshape(-1, 3).copy()
        self.velocity = _f64(velocity).reshape(-1, 3).copy()
        if self.position.shape != self.velocity.shape or self.position.shape[0] == 0:
            # the reference indexes particle_type[0] and panics on an empty type (integrator.rs:29)
            raise MdError(1, "State needs the same non-zero number of positions and velocities")
        self.n = self.position.shape[0]
        self.mass, self.radius, self.id = float(mass), float(radius), int(particle_id)
        self.boundary_box = _f64(boundary_box).reshape(3).copy()
        self.force = np.zeros_like(self.position)      # Particle.force
        self.potential = np.zeros(self.n)              # Particle.potential
        self.temp = np.zeros(self.n)                   # Particle.temp (Σ F_ij · r_ij)

    @staticmethod
    def from_particles(ids, position, velocity, masses, boundary_box):
        if len(set(int(i) for i in ids)) != 1:
            return MultiState.from_particles(ids, position, velocity, masses, boundary_box)
        return State(position, velocity, masses[int(ids[0])], boundary_box, particle_id=int(ids[0]))


class MultiState:
    """core::State with several particle types (particle.rs:24-32): `particles[t]` for t = 0..T-1 stored one after the
    other (`counts[t]` atoms of mass `masses[t]`), as Solver.upload_typed takes them."""

    def __init__(self, position, velocity, counts, masses, boundary_box):
        self.position = _f64(position).reshape(-1, 3).copy()
        self.velocity = _f64(velocity).reshape(-1, 3).copy()
        self.counts = np.asarray(counts, dtype=np.int64).copy()
        self.masses = _f64(masses).copy()
        self.n = self.position.shape[0]
        if self.position.shape != self.velocity.shape or self.counts.sum() != self.n or (self.counts <= 0).any() \
                or self.masses.shape != self.counts.shape:
            # the reference indexes particle_type[0] and panics on an empty type (integrator.rs:29)
            raise MdError(1, "MultiState: every type needs at least one atom; counts must add up to the atom count")
        self.boundary_box = _f64(boundary_box).reshape(3).copy()
        self.force = np.zeros_like(self.position)
        self.potential = np.zeros(self.n)
        self.temp = np.zeros(self.n)

    @staticmethod
    def from_particles(ids, position, velocity, masses, boundary_box):
        """Particles in any order with their type ids → grouped by type id, order kept inside a type (save_data.rs:129-151
        buckets loaded particles the same way)."""
        ids = np.asarray(ids, dtype=np.int64)
        T = int(ids.max()) + 1
        order = np.argsort(ids, kind="stable")
        counts = np.bincount(ids, minlength=T)
        return MultiState(_f64(position).reshape(-1, 3)[order], _f64(velocity).reshape(-1, 3)[order], counts,
                          [masses[t] for t in range(T)], boundary_box)


# --------------------------------------------------------------------------------------------------
class Solver:
    """Device-resident session: owns an md_ctx. `exact=True` selects MD_FORCE_EXACT (bit-identical forces)."""

    def __init__(self, device=0, exact=False, host_loop=False, chunk_loop=False, skin=0.0, max_neighbours=0, cell_subdiv=0,
                 cell_atoms=0.0):
        """host_loop: one host round trip per step (MD_LOOP_HOST); chunk_loop: the graph-chunk loop of the two-kernel step
        for every system (MD_LOOP_CHUNK) instead of the persistent step loop for dilute ones."""
        self._ctx = C.c_void_p()
        cfg = _ffi.Config(device, _ffi.FORCE_EXACT if exact else _ffi.FORCE_FAST,
                          _ffi.LOOP_HOST if host_loop else (_ffi.LOOP_CHUNK if chunk_loop else _ffi.LOOP_AUTO),
                          max_neighbours, cell_subdiv, 0, skin, cell_atoms)
        L = _ffi.lib()
        rc = L.md_create(C.byref(cfg), C.byref(self._ctx))
        if rc != _ffi.MD_OK:
            msg = L.md_last_error(None)
            self._ctx = C.c_void_p()
            raise MdError(rc, msg.decode() if msg else "")
        self.n = 0

    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            _ffi.lib().md_destroy(self._ctx)
            self._ctx = C.c_void_p()

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _ck(self, rc):
        _ffi.check(self._ctx, rc)

    # -- potential / state -------------------------------------------------------------------------
    def set_potential(self, p: Potential):
        self._ck(_ffi.lib().md_set_potential_lj(self._ctx, p.sigma, p.eps, p.r_cut, p.u_cut))

    def set_potential_pair(self, id0, id1, p: Potential):
        """PotentialsDatabase::set_potential(id0, id1, p) (potential.rs:141-144)."""
        self._ck(_ffi.lib().md_set_potential_pair(self._ctx, int(id0), int(id1), p.sigma, p.eps, p.r_cut, p.u_cut))

    def set_cross_type_mode(self, symmetric: bool):
        """False (default): the reference's update_force — a type only accumulates partners of types with the same or a
        larger id (potential.rs:168-176).  True: every pair acts on both atoms."""
        self._ck(_ffi.lib().md_set_cross_type_mode(self._ctx, _ffi.CROSS_SYMMETRIC if symmetric else _ffi.CROSS_REFERENCE))

    def upload_typed(self, state: "MultiState", with_forces=True):
        f = _ptr(state.force) if with_forces else None
        u = _ptr(state.potential) if with_forces else None
        w = _ptr(state.temp) if with_forces else None
        counts = np.ascontiguousarray(state.counts, dtype=np.int64)
        masses = _f64(state.masses)
        box = _f64(state.boundary_box)
        self._ck(_ffi.lib().md_upload_state_typed(self._ctx, state.n, _ptr(state.position), _ptr(state.velocity), f, u, w,
                                                  len(counts), _ptr(counts), _ptr(masses), _ptr(box)))
        self.n = state.n

    def macro_type(self, type_id):
        """The reference's macro parameters of one particle type (macro_parameters/*.rs take a particle_type_id)."""
        m = _ffi.MacroOut()
        self._ck(_ffi.lib().md_macro_type(self._ctx, int(type_id), C.byref(m)))
        return {"kinetic": m.kinetic_energy, "thermal": m.thermal_energy, "potential": m.potential_energy,
                "temperature": m.temperature, "pressure": m.pressure, "vcom": np.array(m.vcom[:]),
                "momentum": np.array(m.momentum[:]), "box": np.array(m.box[:]), "lambda": m.lambda_,
                "myu": m.myu, "n": m.n}

    def upload(self, state: State, with_forces=True):
        f = state.force if with_forces else None
        u = state.potential if with_forces else None
        w = state.temp if with_forces else None
        self.upload_arrays(state.position, state.velocity, state.mass, state.boundary_box, f, u, w)

    def upload_arrays(self, pos, vel, mass, box, force=None, potential=None, virial=None, n=None):
        """Arrays may be numpy arrays or raw host addresses (ints, e.g. pinned torch tensors' data_ptr())."""
        def p(a):
            if a is None or isinstance(a, int):
                return a
            return _ptr(_f64(a))
        if n is None:
            n = np.asarray(pos).size // 3
        box = _f64(box)
        self._ck(_ffi.lib().md_upload_state(self._ctx, n, p(pos), p(vel), p(force), p(potential), p(virial),
                                            float(mass), _ptr(box)))
        self.n = n

    def download(self, state: State):
        box = np.zeros(3)
        self._ck(_ffi.lib().md_download_state(self._ctx, _ptr(state.position), _ptr(state.velocity),
                                              _ptr(state.force), _ptr(state.potential), _ptr(state.temp), _ptr(box)))
        state.boundary_box[:] = box

    def download_arrays(self, pos=None, vel=None, force=None, potential=None, virial=None):
        def p(a):
            return a if (a is None or isinstance(a, int)) else _ptr(a)
        box = np.zeros(3)
        self._ck(_ffi.lib().md_download_state(self._ctx, p(pos), p(vel), p(force), p(potential), p(virial), _ptr(box)))
        return box

    # -- hot path ----------------------------------------------------------------------------------
    def update_force(self):
        self._ck(_ffi.lib().md_update_force(self._ctx))

    def initialize_lattice(self, size, unit_cell, mass, temperature, cell="u", start=None, seed=0):
        """`moldyn_cli initialize` on the device (position.rs:24-104, velocity.rs:6-29): no host State, no upload."""
        sz = np.ascontiguousarray(size, dtype=np.int32).reshape(3)
        st = None if start is None else _f64(start).reshape(3)
        kind = {"u": _ffi.CELL_UNIFORM, "fcc": _ffi.CELL_FCC}[cell]
        self._ck(_ffi.lib().md_initialize_lattice(self._ctx, kind, _ptr(sz), _ptr(st) if st is not None else None,
                                                  float(unit_cell), float(mass), float(temperature), int(seed)))
        self.n = int(sz.prod()) * (4 if cell == "fcc" else 1)

    def step(self, n_steps, dt, thermostat=None, barostat=None):
        """n_steps × Integrator::calculate. thermostat=(Thermostat, target_K), barostat=(Barostat, target_P)."""
        th = thermostat[0]._c(thermostat[1]) if thermostat else None
        ba = barostat[0]._c(barostat[1]) if barostat else None
        self._ck(_ffi.lib().md_step(self._ctx, int(n_steps), float(dt), C.byref(th) if th else None,
                                    C.byref(ba) if ba else None))
        if thermostat:
            thermostat[0].lambda_, thermostat[0].psi = th.lambda_, th.psi
        if barostat:
            barostat[0].myu = ba.myu

    def time_kernels(self, n_steps, dt, thermostat=None, barostat=None):
        """Same as step(), host-stepped with CUDA events around each kernel → {name: (ms_total, launches)}."""
        th = thermostat[0]._c(thermostat[1]) if thermostat else None
        ba = barostat[0]._c(barostat[1]) if barostat else None
        ms = np.zeros(4)
        cnt = np.zeros(4, dtype=np.int64)
        self._ck(_ffi.lib().md_time_kernels(self._ctx, int(n_steps), float(dt), C.byref(th) if th else None,
                                            C.byref(ba) if ba else None, _ptr(ms), _ptr(cnt)))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(("kick_drift", "force", "rebuild", "loop_barrier"))}

    def macro(self):
        m = _ffi.MacroOut()
        self._ck(_ffi.lib().md_macro(self._ctx, C.byref(m)))
        return {"kinetic": m.kinetic_energy, "thermal": m.thermal_energy, "potential": m.potential_energy,
                "temperature": m.temperature, "pressure": m.pressure, "vcom": np.array(m.vcom[:]),
                "momentum": np.array(m.momentum[:]), "box": np.array(m.box[:]), "lambda": m.lambda_,
                "myu": m.myu, "n": m.n}

    # -- introspection -----------------------------------------------------------------------------
    def cells(self):
        cell = np.zeros(self.n, dtype=np.int32)
        dims = np.zeros(3, dtype=np.int32)
        self._ck(_ffi.lib().md_download_cells(self._ctx, _ptr(cell), _ptr(dims)))
        return cell, dims

    def neighbour_lists(self):
        counts = np.zeros(self.n, dtype=np.int64)
        self._ck(_ffi.lib().md_neighbour_counts(self._ctx, _ptr(counts)))
        offsets = np.zeros(self.n + 1, dtype=np.int64)
        np.cumsum(counts, out=offsets[1:])
        partners = np.zeros(max(int(offsets[-1]), 1), dtype=np.int64)
        self._ck(_ffi.lib().md_neighbour_lists(self._ctx, _ptr(offsets), _ptr(partners)))
        return offsets, partners[: int(offsets[-1])]

    def stats(self):
        s = _ffi.Stats()
        self._ck(_ffi.lib().md_get_stats(self._ctx, C.byref(s)))
        return {"steps": s.steps, "rebuilds": s.rebuilds, "kernel_launches": s.kernel_launches,
                "graph_launches": s.graph_launches, "loop_launches": s.loop_launches, "loop_steps": s.loop_steps,
                "cells": list(s.cells), "nbr_capacity": s.nbr_capacity, "nbr_max": s.nbr_max, "skin": s.skin,
                "nbr_mean": s.nbr_mean, "n_owned": s.n_owned, "n_ghost": s.n_ghost, "migrated": s.migrated,
                "wait_halo_ms": s.wait_halo_ms, "wait_sums_ms": s.wait_sums_ms, "peer_memory": s.peer_memory,
                "persistent_loop": s.persistent_loop, "tile_lists": s.tile_lists, "force_atoms_ms": s.force_atoms_ms,
                "force_tail_ms": s.force_tail_ms, "rebuild_ms": s.rebuild_ms, "loop_phase_ms": list(s.loop_phase_ms)}

    def measure_fp64_peak(self):
        """TFLOP/s of a register-only DFMA loop on this device (roofline denominator of the dense force kernel)."""
        t = C.c_double()
        self._ck(_ffi.lib().md_measure_fp64_peak(self._ctx, C.byref(t)))
        return t.value

    def stream(self):
        return _ffi.lib().md_stream(self._ctx)

    def synchronize(self):
        self._ck(_ffi.lib().md_synchronize(self._ctx))

    def invalidate_lists(self):
        self._ck(_ffi.lib().md_invalidate_lists(self._ctx))

    # -- multi-GPU (one process per GPU; see moldyn_b200/distributed.py for the torch.distributed plumbing) --------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_uint8 * _ffi.UNIQUE_ID_BYTES)()
        _ffi.check(None, _ffi.lib().md_comm_unique_id(C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def comm_init(self, rank, nranks, unique_id: bytes):
        buf = (C.c_uint8 * _ffi.UNIQUE_ID_BYTES).from_buffer_copy(unique_id)
        self._ck(_ffi.lib().md_comm_init(self._ctx, int(rank), int(nranks), C.cast(buf, C.c_void_p)))
        self.rank, self.nranks = int(rank), int(nranks)

    def local_count(self):
        a, b = C.c_int64(), C.c_int64()
        self._ck(_ffi.lib().md_local_count(self._ctx, C.byref(a), C.byref(b)))
        return a.value, b.value

    def download_local(self):
        """This rank's owned atoms: dict(ids, position, velocity, force, potential, temp, box)."""
        n, _ = self.local_count()
        ids = np.zeros(n, dtype=np.int64)
        pos, vel, force = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros((n, 3))
        pot, vir, box = np.zeros(n), np.zeros(n), np.zeros(3)
        self._ck(_ffi.lib().md_download_local(self._ctx, _ptr(ids), _ptr(pos), _ptr(vel), _ptr(force), _ptr(pot),
                                              _ptr(vir), _ptr(box)))
        return {"ids": ids, "position": pos, "velocity": vel, "force": force, "potential": pot, "temp": vir,
                "box": box}

    # -- one-shot host-buffer forms ----------------------------------------------------------------
    def update_force_host(self, state: State):
        self._ck(_ffi.lib().md_update_force_host(self._ctx, state.n, _ptr(state.position), state.mass,
                                                 _ptr(state.boundary_box), _ptr(state.force),
                                                 _ptr(state.potential), _ptr(state.temp)))

    def calculate_host(self, state: State, dt, thermostat=None, barostat=None):
        th = thermostat[0]._c(thermostat[1]) if thermostat else None
        ba = barostat[0]._c(barostat[1]) if barostat else None
        self._ck(_ffi.lib().md_calculate_host(self._ctx, state.n, _ptr(state.position), _ptr(state.velocity),
                                              _ptr(state.force), _ptr(state.potential), _ptr(state.temp),
                                              state.mass, _ptr(state.boundary_box), float(dt),
                                              C.byref(th) if th else None, C.byref(ba) if ba else None))
        if thermostat:
            thermostat[0].lambda_, thermostat[0].psi = th.lambda_, th.psi
        if barostat:
            barostat[0].myu = ba.myu


# --------------------------------------------------------------------------------------------------
# Free functions with the reference's signatures (per-call semantics: upload → compute → download).
_default_solver = None


def _solver(exact=None):
    global _default_solver
    if _default_solver is None:
        _default_solver = Solver(exact=bool(int(os.environ.get("MOLDYN_B200_EXACT", "0"))))
    return _default_solver


def update_force(potentials_database: PotentialsDatabase, state: State, solver: Solver | None = None):
    """update_force(&PotentialsDatabase, &mut State) — potential.rs:158."""
    s = solver or _solver()
    s.set_potential(potentials_database.get_potential(state.id, state.id))
    s.update_force_host(state)


class Integrator:
    """Integrator enum (integrator.rs:5-10)."""

    def __init__(self, kind, name=None):
        self.kind, self.name = kind, name

    def calculate(self, potentials_database, state, delta_time, barostat=None, thermostat=None, solver=None):
        """Integrator::calculate(&self, &db, &mut State, dt, &mut Option<(&mut Barostat, f64)>,
        &mut Option<(&mut Thermostat, f64)>) — integrator.rs:14-15."""
        if self.kind != "VerletMethod":
            raise MdError(4, "Integrator::Custom is todo!() in the reference")
        s = solver or _solver()
        s.set_potential(potentials_database.get_potential(state.id, state.id))
        s.calculate_host(state, delta_time, thermostat=thermostat, barostat=barostat)


Integrator.VerletMethod = Integrator("VerletMethod")
Integrator.Custom = staticmethod(lambda name: Integrator("Custom", name))


def _macro_of(state: State, solver=None):
    s = solver or _solver()
    s.upload(state, with_forces=True)
    return s.macro()


def get_center_of_mass_velocity(state, particle_type_id=0, solver=None):  # macro_parameters/mod.rs:12-25
    return _macro_of(state, solver)["vcom"]


def get_momentum_of_system(state, particle_type_id=0, solver=None):  # mod.rs:28-34
    return _macro_of(state, solver)["momentum"]


def get_kinetic_energy(state, particle_type_id=0, solver=None):  # energy.rs:14-22
    return _macro_of(state, solver)["kinetic"]


def get_thermal_energy(state, particle_type_id=0, center_of_mass_velocity=None, solver=None):  # energy.rs:25-37
    return _macro_of(state, solver)["thermal"]


def get_potential_energy(state, particle_type_id=0, solver=None):  # energy.rs:40-49
    return _macro_of(state, solver)["potential"]


def get_temperature(thermal_energy, number_particles):  # temperature.rs:4-7 (two flops: host scalar)
    return (2.0 * thermal_energy) / (3.0 * float(number_particles) * K_B) * 100.0


def get_pressure(state, particle_type_id=0, center_of_mass_velocity=None, solver=None):  # pressure.rs:5-20
    return _macro_of(state, solver)["pressure"]

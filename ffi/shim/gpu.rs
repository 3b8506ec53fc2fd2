//! `solver/src/solver/gpu.rs` — the shim a maintainer adds to the solver crate (next to `potential.rs` / `integrator.rs`):
//! the reference's own signatures, forwarding to libmoldyn_b200.so through `moldyn-b200-sys`.
//!
//! * `update_force(&PotentialsDatabase, &mut State)`            replaces solver/src/solver/potential.rs:158-216
//! * `Integrator::calculate(..)` → `calculate_gpu(..)`          replaces solver/src/solver/integrator.rs:14-59
//! * `GpuSession`                                               device-resident form for `moldyn_cli solve`
//!                                                              (cli/src/commands.rs:102,185)
//!
//! `State` is array-of-structs (`Vec<Vec<Particle>>`, core/src/particle.rs:6-32); the C ABI takes one particle type as
//! flat xyz-interleaved arrays, so `upload`/`download` marshal `state.particles` (type by type) field by field.
use moldyn_b200_sys as sys;
use moldyn_core::{Particle, State};
use nalgebra::Vector3;
use std::cell::RefCell;
use std::ffi::CStr;

use crate::initializer::{Barostat, Thermostat};
use crate::solver::{Integrator, Potential, PotentialsDatabase};

pub struct GpuSession {
    ctx: *mut sys::md_ctx,
    pos: Vec<f64>,
    vel: Vec<f64>,
    force: Vec<f64>,
    pot: Vec<f64>,
    vir: Vec<f64>,
}

/// A non-zero md_status becomes the panic the reference raises at the same point (`expect` / `todo!()`).
fn check(ctx: *const sys::md_ctx, rc: i32) {
    if rc != sys::MD_OK {
        let msg = unsafe { CStr::from_ptr(sys::md_last_error(ctx)) }.to_string_lossy().into_owned();
        if rc == sys::MD_ERR_UNSUPPORTED {
            todo!("{msg}");
        }
        panic!("moldyn_b200 error {rc}: {msg}");
    }
}

impl GpuSession {
    pub fn new() -> Self {
        let cfg = sys::md_config::default(); // device 0, MD_FORCE_FAST, MD_LOOP_AUTO
        let mut ctx = std::ptr::null_mut();
        check(std::ptr::null(), unsafe { sys::md_create(&cfg, &mut ctx) });
        GpuSession { ctx, pos: vec![], vel: vec![], force: vec![], pot: vec![], vir: vec![] }
    }

    pub fn set_potential(&mut self, potential: &Potential) {
        match potential {
            Potential::LennardJones { sigma, eps, r_cut, u_cut } => {
                check(self.ctx, unsafe { sys::md_set_potential_lj(self.ctx, *sigma, *eps, *r_cut, *u_cut) })
            }
            Potential::Custom { .. } => todo!(), // potential.rs:71-73
        }
    }

    /// State → device (core/src/particle.rs:6-32).  `with_forces = false` is the State right after loading a frame
    /// (save_data.rs:86-98: force, potential and temp are zero).
    pub fn upload(&mut self, state: &State, with_forces: bool) {
        // State.particles is indexed by type id: the device takes the types one after the other (md_upload_state_typed); one
        // type is the plain md_upload_state.  A type without atoms panics in the reference (particle_type[0], integrator.rs:29).
        assert!(!state.particles.is_empty() && state.particles.iter().all(|t| !t.is_empty()), "every particle type needs an atom");
        let n: usize = state.particles.iter().map(|t| t.len()).sum();
        self.pos.clear();
        self.vel.clear();
        self.force.clear();
        self.pot.clear();
        self.vir.clear();
        for p in state.particles.iter().flatten() {
            self.pos.extend_from_slice(&[p.position.x, p.position.y, p.position.z]);
            self.vel.extend_from_slice(&[p.velocity.x, p.velocity.y, p.velocity.z]);
            self.force.extend_from_slice(&[p.force.x, p.force.y, p.force.z]);
            self.pot.push(p.potential);
            self.vir.push(p.temp);
        }
        let counts: Vec<i64> = state.particles.iter().map(|t| t.len() as i64).collect();
        let masses: Vec<f64> = state.particles.iter().map(|t| t[0].mass).collect(); // integrator.rs:30
        let bb = [state.boundary_box.x, state.boundary_box.y, state.boundary_box.z];
        let null = std::ptr::null::<f64>();
        check(self.ctx, unsafe {
            sys::md_upload_state_typed(
                self.ctx, n as i64, self.pos.as_ptr(), self.vel.as_ptr(),
                if with_forces { self.force.as_ptr() } else { null },
                if with_forces { self.pot.as_ptr() } else { null },
                if with_forces { self.vir.as_ptr() } else { null },
                counts.len() as i32, counts.as_ptr(), masses.as_ptr(), bb.as_ptr(),
            )
        });
    }

    /// Every potential the State's particle types can meet: get_potential(a, b) for a <= b < T (the (min, max) key or the
    /// default, potential.rs:147-155).  One type: its own pair only.
    pub fn set_potentials(&mut self, potentials_database: &PotentialsDatabase, state: &State) {
        let n_types = state.particles.len() as u16;
        for a in 0..n_types {
            for b in a..n_types {
                match potentials_database.get_potential(a, b) {
                    Potential::LennardJones { sigma, eps, r_cut, u_cut } => check(self.ctx, unsafe {
                        sys::md_set_potential_pair(self.ctx, a as i32, b as i32, *sigma, *eps, *r_cut, *u_cut)
                    }),
                    Potential::Custom { .. } => todo!(), // potential.rs:71-73
                }
            }
        }
    }

    /// false (default): the reference's one-sided cross-type accumulation (potential.rs:168-176); true: symmetric table.
    pub fn set_symmetric_cross_type_forces(&mut self, symmetric: bool) {
        let mode = if symmetric { sys::MD_CROSS_SYMMETRIC } else { sys::MD_CROSS_REFERENCE };
        check(self.ctx, unsafe { sys::md_set_cross_type_mode(self.ctx, mode) });
    }

    /// device → State: positions, velocities, forces, potential, temp (= Σ F·r) and boundary_box, in upload order.
    pub fn download(&mut self, state: &mut State) {
        let n: usize = state.particles.iter().map(|t| t.len()).sum();
        self.pos.resize(3 * n, 0.0);
        self.vel.resize(3 * n, 0.0);
        self.force.resize(3 * n, 0.0);
        self.pot.resize(n, 0.0);
        self.vir.resize(n, 0.0);
        let mut bb = [0.0f64; 3];
        check(self.ctx, unsafe {
            sys::md_download_state(self.ctx, self.pos.as_mut_ptr(), self.vel.as_mut_ptr(), self.force.as_mut_ptr(),
                                   self.pot.as_mut_ptr(), self.vir.as_mut_ptr(), bb.as_mut_ptr())
        });
        for (i, p) in state.particles.iter_mut().flatten().enumerate() {
            p.position = Vector3::new(self.pos[3 * i], self.pos[3 * i + 1], self.pos[3 * i + 2]);
            p.velocity = Vector3::new(self.vel[3 * i], self.vel[3 * i + 1], self.vel[3 * i + 2]);
            p.force = Vector3::new(self.force[3 * i], self.force[3 * i + 1], self.force[3 * i + 2]);
            p.potential = self.pot[i];
            p.temp = self.vir[i];
        }
        state.boundary_box = Vector3::new(bb[0], bb[1], bb[2]);
    }

    pub fn update_force(&mut self) {
        check(self.ctx, unsafe { sys::md_update_force(self.ctx) });
    }

    /// n × Integrator::VerletMethod.calculate on the resident state; lambda / psi / myu are stored back into the enums
    /// like thermostat.rs:33,37-38 and barostat.rs:30 do.
    pub fn step(&mut self, n_steps: i64, delta_time: f64, barostat: &mut Option<(&mut Barostat, f64)>,
                thermostat: &mut Option<(&mut Thermostat, f64)>) {
        let mut th = thermostat.as_ref().map(|(t, target)| thermostat_to_c(t, *target));
        let mut ba = barostat.as_ref().map(|(b, target)| barostat_to_c(b, *target));
        let th_ptr = th.as_mut().map_or(std::ptr::null_mut(), |t| t as *mut sys::md_thermostat);
        let ba_ptr = ba.as_mut().map_or(std::ptr::null_mut(), |b| b as *mut sys::md_barostat);
        check(self.ctx, unsafe { sys::md_step(self.ctx, n_steps, delta_time, th_ptr, ba_ptr) });
        if let (Some((t, _)), Some(c)) = (thermostat.as_mut(), th) {
            match t {
                Thermostat::Berendsen { lambda, .. } => *lambda = c.lambda,
                Thermostat::NoseHoover { psi, lambda, .. } => { *psi = c.psi; *lambda = c.lambda; }
                Thermostat::Custom { .. } => {}
            }
        }
        if let (Some((b, _)), Some(c)) = (barostat.as_mut(), ba) {
            if let Barostat::Berendsen { myu, .. } = b { *myu = c.myu; }
        }
    }
}

impl Drop for GpuSession {
    fn drop(&mut self) {
        unsafe { sys::md_destroy(self.ctx) };
    }
}

fn thermostat_to_c(t: &Thermostat, target: f64) -> sys::md_thermostat {
    match t {
        Thermostat::Berendsen { tau, lambda } => sys::md_thermostat {
            kind: sys::MD_THERMOSTAT_BERENDSEN, reserved0: 0, tau: *tau, target, lambda: *lambda, psi: 0.0 },
        Thermostat::NoseHoover { tau, psi, lambda } => sys::md_thermostat {
            kind: sys::MD_THERMOSTAT_NOSE_HOOVER, reserved0: 0, tau: *tau, target, lambda: *lambda, psi: *psi },
        Thermostat::Custom { .. } => sys::md_thermostat { kind: 99, ..Default::default() }, // → MD_ERR_UNSUPPORTED → todo!()
    }
}

fn barostat_to_c(b: &Barostat, target: f64) -> sys::md_barostat {
    match b {
        Barostat::Berendsen { beta, tau, myu } => sys::md_barostat {
            kind: sys::MD_BAROSTAT_BERENDSEN, reserved0: 0, beta: *beta, tau: *tau, target, myu: *myu },
        Barostat::Custom { .. } => sys::md_barostat { kind: 99, ..Default::default() },
    }
}

thread_local! {
    // the free-function API of the reference has no session argument: one per calling thread
    static SESSION: RefCell<Option<GpuSession>> = RefCell::new(None);
}

fn with_session<R>(f: impl FnOnce(&mut GpuSession) -> R) -> R {
    SESSION.with(|s| f(s.borrow_mut().get_or_insert_with(GpuSession::new)))
}

/// Drop-in for `solver::update_force` (potential.rs:158-216): per-call semantics, State in, State out.
pub fn update_force(potentials_database: &PotentialsDatabase, state: &mut State) {
    with_session(|s| {
        s.set_potentials(potentials_database, state);
        s.upload(state, false);
        s.update_force();
        s.download(state);
    })
}

/// Drop-in body of `Integrator::calculate` (integrator.rs:14-59), per-call semantics.  `moldyn_cli solve` should keep a
/// `GpuSession` instead and download only at frame boundaries.
pub fn calculate_gpu(integrator: &Integrator, potentials_database: &PotentialsDatabase, state: &mut State, delta_time: f64,
                     barostat: &mut Option<(&mut Barostat, f64)>, thermostat: &mut Option<(&mut Thermostat, f64)>) {
    let Integrator::VerletMethod = integrator else { todo!() }; // integrator.rs:60-62
    with_session(|s| {
        s.set_potentials(potentials_database, state);
        s.upload(state, true);
        s.step(1, delta_time, barostat, thermostat);
        s.download(state);
    })
}

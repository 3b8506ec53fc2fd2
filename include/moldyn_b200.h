/*
 * moldyn_b200.h — C ABI of the B200-native replacement for moldyn's `solve` step loop.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  Every
 * entry point names the reference interface it replaces (paths relative to the
 * AndrewChe7/moldyn repository).  A Rust shim keeps the reference's
 * enum/method signatures and forwards to these functions.
 *
 * Model: a context owns a device-resident copy of ONE particle type's State
 * (core/src/particle.rs:6-32) in f64.  The reference's per-call semantics
 * (`update_force(&db, &mut state)`, `Integrator::calculate(..)`) are available both as
 * a device-resident session (upload once, step many times, download at frame
 * boundaries) and as one-shot host-buffer calls (md_update_force_host, md_calculate_host).
 *
 * Conventions: every function returns an md_status (0 = ok) and never unwinds; the
 * message of the last failure is md_last_error(ctx) (md_last_error(NULL) for
 * md_create failures).  Host arrays are borrowed only for the duration of the call.
 * Vectors are xyz-interleaved (pos[3*i+d]), like Vec<Vector3<f64>>.  A context is
 * single-owner and not thread-safe (the reference solver is called from one thread).
 * There is no CPU fallback: without a CUDA device md_create fails with MD_ERR_CUDA.
 */
#ifndef MOLDYN_B200_H
#define MOLDYN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MD_API __attribute__((visibility("default")))

typedef enum md_status {
    MD_OK = 0,
    MD_ERR_INVALID_ARGUMENT = 1,
    MD_ERR_CUDA = 2,
    MD_ERR_NCCL = 3,
    MD_ERR_UNSUPPORTED = 4,        /* Custom variants (todo!() in the reference), multi-type states on several GPUs */
    MD_ERR_NEIGHBOUR_OVERFLOW = 5, /* neighbour list could not be grown */
    MD_ERR_NO_STATE = 6,           /* step/download before md_upload_state */
    MD_ERR_NONFINITE = 7,          /* NaN/inf reached the step controls (e.g. Berendsen lambda at T = 0) */
    MD_ERR_DECOMPOSITION = 8       /* box too small for the requested number of ranks */
} md_status;

/* force_mode */
#define MD_FORCE_FAST 0  /* r^2-based LJ, FMA allowed, neighbour order = cell order (default) */
#define MD_FORCE_EXACT 1 /* the reference's operation order (potential.rs:181-211) without FMA and with
                            partners summed in ascending particle index: bit-identical to update_force */
/* loop_mode */
#define MD_LOOP_AUTO 0  /* default.  Dilute systems (mean listed partners < 8): ONE persistent cooperative kernel runs the steps
                           until the list must be rebuilt (k_md_loop: two grid-wide synchronisations per step, no kernel boundary).
                           Dense systems: pre-enqueued CUDA graphs of 16 guarded {k_kick_drift; k_force} steps (MD_LOOP_CHUNK) */
#define MD_LOOP_HOST 1  /* one host round-trip per step (debugging / cross-check, ncu): same kernels, same bits as MD_LOOP_AUTO */
#define MD_LOOP_CHUNK 2 /* the graph-chunk loop of the two-kernel step for every system (the round-1 path; A/B measurements) */

typedef struct md_config {
    int32_t device;         /* CUDA device ordinal */
    int32_t force_mode;     /* MD_FORCE_* */
    int32_t loop_mode;      /* MD_LOOP_* */
    int32_t max_neighbours; /* initial neighbour-list capacity per atom; 0 = from density (grown on demand) */
    int32_t cell_subdiv;    /* cells per (r_cut+skin): 1 (27-cell stencil) or 2 (125-cell stencil); 0 = by density */
    int32_t reserved0;
    double skin;            /* Verlet skin [nm]; <= 0 selects a default from r_cut and density */
    double cell_atoms;      /* target atoms per cell for dilute systems; <= 0 = 1 */
} md_config;

/* Thermostat (solver/src/initializer/thermostat.rs:4-22).  lambda/psi are written back after
 * md_step, like the reference stores them in the enum (thermostat.rs:33,37-38). */
#define MD_THERMOSTAT_NONE 0
#define MD_THERMOSTAT_BERENDSEN 1
#define MD_THERMOSTAT_NOSE_HOOVER 2 /* thermostat.rs:35-39,59-65: psi is carried by the caller, like the enum field */
typedef struct md_thermostat {
    int32_t kind;
    int32_t reserved0;
    double tau;
    double target; /* target temperature [K] — the f64 paired with the thermostat in Integrator::calculate */
    double lambda; /* out */
    double psi;    /* in/out (Nose-Hoover) */
} md_thermostat;

/* Barostat (solver/src/initializer/barostat.rs:4-19). */
#define MD_BAROSTAT_NONE 0
#define MD_BAROSTAT_BERENDSEN 1
typedef struct md_barostat {
    int32_t kind;
    int32_t reserved0;
    double beta;
    double tau;
    double target; /* target pressure — the f64 paired with the barostat in Integrator::calculate */
    double myu;    /* out */
} md_barostat;

/* macro_parameters::* of the resident state (solver/src/macro_parameters/{mod,energy,temperature,pressure}.rs) */
typedef struct md_macro_out {
    double kinetic_energy;   /* get_kinetic_energy   energy.rs:14-22 */
    double thermal_energy;   /* get_thermal_energy   energy.rs:25-37 */
    double potential_energy; /* get_potential_energy energy.rs:40-49 */
    double temperature;      /* get_temperature [K]  temperature.rs:4-7 */
    double pressure;         /* get_pressure         pressure.rs:5-20 (virial of the last force evaluation) */
    double vcom[3];          /* get_center_of_mass_velocity mod.rs:12-25 */
    double momentum[3];      /* get_momentum_of_system      mod.rs:28-34 */
    double box[3];           /* State::boundary_box */
    double lambda;           /* last Berendsen lambda (1 if none) */
    double myu;              /* last Berendsen myu (1 if none) */
    int64_t n;
} md_macro_out;

typedef struct md_stats {
    int64_t steps;            /* MD steps executed since md_create */
    int64_t rebuilds;         /* neighbour-list rebuilds */
    int64_t kernel_launches;  /* kernels of this library launched (graph body launches included) */
    int64_t graph_launches;   /* step-chunk graph launches */
    int64_t loop_launches;    /* launches of the persistent step loop (k_md_loop) */
    int64_t loop_steps;       /* steps executed inside the persistent step loop since the last upload */
    int32_t cells[3];         /* current cell grid */
    int32_t nbr_capacity;     /* neighbour slots per atom */
    int32_t nbr_max;          /* largest neighbour count seen at the last rebuild */
    int32_t peer_memory;      /* 1: halo and reduction go through peer memory (NVLink stores), 0: NCCL send/recv path */
    int32_t persistent_loop;  /* 1: md_step runs this state through the persistent step loop */
    int32_t tile_lists;       /* 1: the last rebuild produced the brick-local lists of the tile force kernel (dense systems) */
    double skin;              /* skin in use */
    double nbr_mean;          /* mean neighbour count at the last rebuild */
    int64_t n_owned;          /* atoms this rank owns (== n on one GPU) */
    int64_t n_ghost;          /* halo atoms held for the neighbours' partners */
    int64_t migrated;         /* atoms handed to a neighbouring rank so far */
    double wait_halo_ms;      /* multi-GPU peer-memory path: time spent polling for the neighbours' ghosts,          */
    double wait_sums_ms;      /* and for the other ranks' reduction sums (since the last upload)                   */
    double force_atoms_ms;    /* two-kernel peer-memory path diagnostics: k_force first block start -> all atoms done, */
    double force_tail_ms;     /*   -> mailbox exchange + finalize done                                             */
    double rebuild_ms;        /* multi-GPU: wall time spent in list rebuilds so far (host clock around dist_rebuild) */
    double loop_phase_ms[4];  /* persistent step loop, block 0's clock, since the last upload: drift phase, mid-step barrier,
                                 force phase, reduction + exchange + finalize + end-of-step wait */
} md_stats;

typedef struct md_ctx md_ctx;

/* ---- lifetime ----------------------------------------------------------------------------- */
MD_API int md_create(const md_config *cfg, md_ctx **out);
MD_API void md_destroy(md_ctx *ctx);
MD_API const char *md_last_error(const md_ctx *ctx);
MD_API const char *md_version(void);

/* ---- Potential (solver/src/solver/potential.rs:12-87) ------------------------------------- */
/* Potential::new_lennard_jones (potential.rs:27-55): r_cut = 2.5 sigma, u_cut = U(r_cut). Host arithmetic. */
MD_API int md_lj_new(double sigma, double eps, double *r_cut, double *u_cut);
/* Potential::get_potential_and_force (potential.rs:57-70). Host arithmetic, scalar. */
MD_API int md_lj_potential_and_force(double sigma, double eps, double r_cut, double u_cut, double r,
                                     double *potential, double *force);
/* PotentialsDatabase::set_potential(0,0,LennardJones{..}) (potential.rs:141-144) for the resident type.
 * Default after md_create: PotentialsDatabase::new() → argon sigma=0.3418 eps=1.712 (potential.rs:95-101). */
MD_API int md_set_potential_lj(md_ctx *ctx, double sigma, double eps, double r_cut, double u_cut);

/* PotentialsDatabase::set_potential(id0, id1, LennardJones{..}) (potential.rs:141-144): the entry keyed (min, max).  Pairs
 * never set fall back to the default potential, argon (potential.rs:147-155).  (0, 0) is md_set_potential_lj. */
#define MD_MAX_TYPES 8
MD_API int md_set_potential_pair(md_ctx *ctx, int32_t id0, int32_t id1, double sigma, double eps, double r_cut,
                                 double u_cut);
/* What update_force does between particle types.  MD_CROSS_REFERENCE (default) is the reference, literally
 * (potential.rs:168-176, `for particle_type2 in particle_type1..`): atoms of type t1 accumulate their partners of types
 * t2 >= t1 only — a type never feels a type with a smaller id.  MD_CROSS_SYMMETRIC lets every atom accumulate its partners
 * of every type: the symmetric type-pair table.  Everything else the reference does with several types is kept in both
 * modes: myu and lambda of the LAST type are applied to all (integrator.rs:18-27), kicks use the type's mass
 * (integrator.rs:29-30), the box is scaled once per type (integrator.rs:54-58). */
#define MD_CROSS_REFERENCE 0
#define MD_CROSS_SYMMETRIC 1
MD_API int md_set_cross_type_mode(md_ctx *ctx, int32_t mode);

/* ---- State transfer (core/src/particle.rs:6-32, core/src/save_data.rs:129-151) ------------ */
/* pos, vel: 3n doubles.  force (3n), potential (n), virial (n = Particle.temp) may be NULL → zero, as after
 * StateToSave → State (save_data.rs:86-98).  mass: ParticleDatabase mass of the type. box: boundary_box. */
MD_API int md_upload_state(md_ctx *ctx, int64_t n, const double *pos, const double *vel, const double *force,
                           const double *potential, const double *virial, double mass, const double box[3]);
/* State with several particle types: State.particles (Vec<Vec<Particle>>, indexed by type id, core/src/particle.rs:24-32)
 * flattened type by type — type t owns type_counts[t] consecutive atoms, type_mass[t] is their Particle.mass.  Every type
 * needs at least one atom (the reference indexes particle_type[0], integrator.rs:29).  md_update_force, md_step (thermostat
 * None / Berendsen, barostat None / Berendsen), md_download_state and md_macro_type then follow the reference's multi-type
 * behaviour (see md_set_cross_type_mode); one GPU; the step is host-stepped (a correctness path, not the tuned one). */
MD_API int md_upload_state_typed(md_ctx *ctx, int64_t n, const double *pos, const double *vel, const double *force,
                                 const double *potential, const double *virial, int32_t n_types,
                                 const int64_t *type_counts, const double *type_mass, const double box[3]);
/* Any output pointer may be NULL. Particles come back in the order they were uploaded. */
MD_API int md_download_state(md_ctx *ctx, double *pos, double *vel, double *force, double *potential,
                             double *virial, double box[3]);

/* Device-side initializer: builds the State on the GPU instead of uploading it.
 * UnitCell::{U,FCC}.initialize_particles_position (solver/src/initializer/position.rs:24-104; positions bit-identical)
 * + initialize_velocities_maxwell_boltzmann (velocity.rs:6-29; the reference's RNG is unseeded, so the velocities match in
 * distribution: first half sigma*N(0,1), second half the negated copy).  boundary_box = unit_cell * size, like
 * `moldyn_cli initialize` (cli/src/commands.rs:43-80).  start may be NULL (= origin). */
#define MD_CELL_UNIFORM 0
#define MD_CELL_FCC 1
MD_API int md_initialize_lattice(md_ctx *ctx, int cell_type, const int32_t size[3], const double start[3],
                                 double unit_cell, double mass, double temperature, uint64_t seed);

/* ---- the hot path ------------------------------------------------------------------------- */
/* update_force(&potentials_db, &mut state) (potential.rs:158-216) on the resident state. */
MD_API int md_update_force(md_ctx *ctx);
/* n_steps × Integrator::VerletMethod.calculate(&db, &mut state, dt, &mut barostat, &mut thermostat)
 * (solver/src/solver/integrator.rs:14-59).  thermostat / barostat may be NULL (= None). */
MD_API int md_step(md_ctx *ctx, int64_t n_steps, double dt, md_thermostat *thermostat, md_barostat *barostat);
/* macro parameters of the resident state. */
MD_API int md_macro(md_ctx *ctx, md_macro_out *out);
/* The same for one particle type of a multi-type State — every macro_parameters function takes a particle_type_id. */
MD_API int md_macro_type(md_ctx *ctx, int32_t type_id, md_macro_out *out);

/* One-shot host-buffer forms with the reference's per-call semantics (all arrays in/out, caller-owned):
 * md_update_force_host ≡ update_force; md_calculate_host ≡ Integrator::calculate (one step). */
MD_API int md_update_force_host(md_ctx *ctx, int64_t n, const double *pos, double mass, const double box[3],
                                double *force, double *potential, double *virial);
MD_API int md_calculate_host(md_ctx *ctx, int64_t n, double *pos, double *vel, double *force, double *potential,
                             double *virial, double mass, double box[3], double dt, md_thermostat *thermostat,
                             md_barostat *barostat);

/* ---- multi-GPU: one process per GPU, 1-D slab decomposition along x ----------------------
 * The reference has no distributed mode (rayon threads only); these entry points have no reference counterpart.
 * Rank 0 creates an id (ncclGetUniqueId), the host layer broadcasts it (torch.distributed / MPI / a file), every
 * rank calls md_comm_init BEFORE md_upload_state.  md_upload_state then takes the FULL state on every rank and
 * keeps the atoms whose fractional x lies in this rank's slab; md_update_force / md_step / md_macro are collective
 * (every rank calls them with the same arguments; md_macro returns the same global values everywhere);
 * md_download_local returns this rank's owned atoms with their upload indices. */
#define MD_UNIQUE_ID_BYTES 128
MD_API int md_comm_unique_id(uint8_t id[MD_UNIQUE_ID_BYTES]);
MD_API int md_comm_init(md_ctx *ctx, int rank, int nranks, const uint8_t id[MD_UNIQUE_ID_BYTES]);
MD_API int md_local_count(md_ctx *ctx, int64_t *n_owned, int64_t *n_ghost);
/* ids[n_owned], pos/vel/force[3 n_owned], potential/virial[n_owned]; any pointer may be NULL */
MD_API int md_download_local(md_ctx *ctx, int64_t *ids, double *pos, double *vel, double *force, double *potential,
                             double *virial, double box[3]);
/* Host-only (no GPU): the slab [x_lo, x_hi) rank owns, its ring neighbours and a per-rank atom capacity hint.
 * MD_ERR_DECOMPOSITION if a slab would be narrower than 2.1 x r_list. */
MD_API int md_plan_decomposition(int64_t n, const double box[3], double r_list, int nranks, int rank, double *x_lo,
                                 double *x_hi, int *left, int *right, int64_t *capacity_hint);

/* ---- introspection used by the parity tests and the bench ---------------------------------- */
/* Cell index of every atom (upload order) at the last list build + grid dims.  cell = (cx*ny + cy)*nz + cz with
 * c_d = min(nc_d-1, (int)(frac(x_d / L_d) * nc_d)). */
MD_API int md_download_cells(md_ctx *ctx, int32_t *cell_of_atom, int32_t dims[3]);
/* Verlet list of the last build in upload indices: counts[i], then partners (ascending) at offsets[i]. */
MD_API int md_neighbour_counts(md_ctx *ctx, int64_t *counts);
MD_API int md_neighbour_lists(md_ctx *ctx, const int64_t *offsets, int64_t *partners);
MD_API int md_get_stats(md_ctx *ctx, md_stats *out);
/* cudaStream_t the context launches on (for CUDA-event timing by the caller). */
MD_API void *md_stream(md_ctx *ctx);
MD_API int md_synchronize(md_ctx *ctx);
/* md_step with device timing of its parts: summed device milliseconds and counts of
 * {k_kick_drift or the loop's drift phase, k_force or the loop's force phase + reduction tail, list rebuild, the loop's
 * mid-step barrier}.  The two-kernel step is host-stepped with a CUDA-event pair around every kernel; the persistent loop
 * reads %globaltimer in block 0.  Measurement aid for the roofline figures. */
MD_API int md_time_kernels(md_ctx *ctx, int64_t n_steps, double dt, md_thermostat *thermostat,
                           md_barostat *barostat, double ms[4], int64_t launches[4]);
/* Measurement aid: the device's sustained FP64 fused-multiply-add rate in TFLOP/s (a register-only DFMA loop on every SM,
 * best of five launches) — the denominator of the dense force kernel's roofline fraction. */
MD_API int md_measure_fp64_peak(md_ctx *ctx, double *tflops);
/* Forces a list rebuild before the next force evaluation (rebuild-stress tests). */
MD_API int md_invalidate_lists(md_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* MOLDYN_B200_H */

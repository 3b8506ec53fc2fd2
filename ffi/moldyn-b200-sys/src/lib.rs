//! Raw bindings of `include/moldyn_b200.h`, one declaration per exported symbol, same order as the header.
//! Every function returns an `md_status` (0 = ok) and never unwinds; `md_last_error` holds the message.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const MD_OK: c_int = 0;
pub const MD_ERR_INVALID_ARGUMENT: c_int = 1;
pub const MD_ERR_CUDA: c_int = 2;
pub const MD_ERR_NCCL: c_int = 3;
pub const MD_ERR_UNSUPPORTED: c_int = 4;
pub const MD_ERR_NEIGHBOUR_OVERFLOW: c_int = 5;
pub const MD_ERR_NO_STATE: c_int = 6;
pub const MD_ERR_NONFINITE: c_int = 7;
pub const MD_ERR_DECOMPOSITION: c_int = 8;

pub const MD_FORCE_FAST: i32 = 0;
pub const MD_FORCE_EXACT: i32 = 1;
pub const MD_LOOP_AUTO: i32 = 0;
pub const MD_LOOP_HOST: i32 = 1;
pub const MD_LOOP_CHUNK: i32 = 2;
pub const MD_THERMOSTAT_NONE: i32 = 0;
pub const MD_THERMOSTAT_BERENDSEN: i32 = 1;
pub const MD_THERMOSTAT_NOSE_HOOVER: i32 = 2;
pub const MD_BAROSTAT_NONE: i32 = 0;
pub const MD_BAROSTAT_BERENDSEN: i32 = 1;
pub const MD_CELL_UNIFORM: c_int = 0;
pub const MD_CELL_FCC: c_int = 1;
pub const MD_UNIQUE_ID_BYTES: usize = 128;
pub const MD_MAX_TYPES: usize = 8;
pub const MD_CROSS_REFERENCE: i32 = 0;
pub const MD_CROSS_SYMMETRIC: i32 = 1;

#[repr(C)]
pub struct md_ctx {
    _private: [u8; 0],
}

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct md_config {
    pub device: i32,
    pub force_mode: i32,
    pub loop_mode: i32,
    pub max_neighbours: i32,
    pub cell_subdiv: i32,
    pub reserved0: i32,
    pub skin: f64,
    pub cell_atoms: f64,
}

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct md_thermostat {
    pub kind: i32,
    pub reserved0: i32,
    pub tau: f64,
    pub target: f64,
    pub lambda: f64,
    pub psi: f64,
}

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct md_barostat {
    pub kind: i32,
    pub reserved0: i32,
    pub beta: f64,
    pub tau: f64,
    pub target: f64,
    pub myu: f64,
}

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct md_macro_out {
    pub kinetic_energy: f64,
    pub thermal_energy: f64,
    pub potential_energy: f64,
    pub temperature: f64,
    pub pressure: f64,
    pub vcom: [f64; 3],
    pub momentum: [f64; 3],
    pub box_: [f64; 3],
    pub lambda: f64,
    pub myu: f64,
    pub n: i64,
}

#[repr(C)]
#[derive(Default, Clone, Copy, Debug)]
pub struct md_stats {
    pub steps: i64,
    pub rebuilds: i64,
    pub kernel_launches: i64,
    pub graph_launches: i64,
    pub loop_launches: i64,
    pub loop_steps: i64,
    pub cells: [i32; 3],
    pub nbr_capacity: i32,
    pub nbr_max: i32,
    pub peer_memory: i32,
    pub persistent_loop: i32,
    pub tile_lists: i32,
    pub skin: f64,
    pub nbr_mean: f64,
    pub n_owned: i64,
    pub n_ghost: i64,
    pub migrated: i64,
    pub wait_halo_ms: f64,
    pub wait_sums_ms: f64,
    pub force_atoms_ms: f64,
    pub force_tail_ms: f64,
    pub rebuild_ms: f64,
    pub loop_phase_ms: [f64; 4],
}

extern "C" {
    pub fn md_create(cfg: *const md_config, out: *mut *mut md_ctx) -> c_int;
    pub fn md_destroy(ctx: *mut md_ctx);
    pub fn md_last_error(ctx: *const md_ctx) -> *const c_char;
    pub fn md_version() -> *const c_char;
    pub fn md_lj_new(sigma: f64, eps: f64, r_cut: *mut f64, u_cut: *mut f64) -> c_int;
    pub fn md_lj_potential_and_force(sigma: f64, eps: f64, r_cut: f64, u_cut: f64, r: f64, potential: *mut f64, force: *mut f64) -> c_int;
    pub fn md_set_potential_lj(ctx: *mut md_ctx, sigma: f64, eps: f64, r_cut: f64, u_cut: f64) -> c_int;
    pub fn md_set_potential_pair(ctx: *mut md_ctx, id0: i32, id1: i32, sigma: f64, eps: f64, r_cut: f64, u_cut: f64) -> c_int;
    pub fn md_set_cross_type_mode(ctx: *mut md_ctx, mode: i32) -> c_int;
    pub fn md_upload_state_typed(ctx: *mut md_ctx, n: i64, pos: *const f64, vel: *const f64, force: *const f64, potential: *const f64, virial: *const f64, n_types: i32, type_counts: *const i64, type_mass: *const f64, box_: *const f64) -> c_int;
    pub fn md_upload_state(ctx: *mut md_ctx, n: i64, pos: *const f64, vel: *const f64, force: *const f64, potential: *const f64, virial: *const f64, mass: f64, box_: *const f64) -> c_int;
    pub fn md_download_state(ctx: *mut md_ctx, pos: *mut f64, vel: *mut f64, force: *mut f64, potential: *mut f64, virial: *mut f64, box_: *mut f64) -> c_int;
    pub fn md_initialize_lattice(ctx: *mut md_ctx, cell_type: c_int, size: *const i32, start: *const f64, unit_cell: f64, mass: f64, temperature: f64, seed: u64) -> c_int;
    pub fn md_update_force(ctx: *mut md_ctx) -> c_int;
    pub fn md_step(ctx: *mut md_ctx, n_steps: i64, dt: f64, thermostat: *mut md_thermostat, barostat: *mut md_barostat) -> c_int;
    pub fn md_macro(ctx: *mut md_ctx, out: *mut md_macro_out) -> c_int;
    pub fn md_macro_type(ctx: *mut md_ctx, type_id: i32, out: *mut md_macro_out) -> c_int;
    pub fn md_update_force_host(ctx: *mut md_ctx, n: i64, pos: *const f64, mass: f64, box_: *const f64, force: *mut f64, potential: *mut f64, virial: *mut f64) -> c_int;
    pub fn md_calculate_host(ctx: *mut md_ctx, n: i64, pos: *mut f64, vel: *mut f64, force: *mut f64, potential: *mut f64, virial: *mut f64, mass: f64, box_: *mut f64, dt: f64, thermostat: *mut md_thermostat, barostat: *mut md_barostat) -> c_int;
    pub fn md_comm_unique_id(id: *mut u8) -> c_int;
    pub fn md_comm_init(ctx: *mut md_ctx, rank: c_int, nranks: c_int, id: *const u8) -> c_int;
    pub fn md_local_count(ctx: *mut md_ctx, n_owned: *mut i64, n_ghost: *mut i64) -> c_int;
    pub fn md_download_local(ctx: *mut md_ctx, ids: *mut i64, pos: *mut f64, vel: *mut f64, force: *mut f64, potential: *mut f64, virial: *mut f64, box_: *mut f64) -> c_int;
    pub fn md_plan_decomposition(n: i64, box_: *const f64, r_list: f64, nranks: c_int, rank: c_int, x_lo: *mut f64, x_hi: *mut f64, left: *mut c_int, right: *mut c_int, capacity_hint: *mut i64) -> c_int;
    pub fn md_download_cells(ctx: *mut md_ctx, cell_of_atom: *mut i32, dims: *mut i32) -> c_int;
    pub fn md_neighbour_counts(ctx: *mut md_ctx, counts: *mut i64) -> c_int;
    pub fn md_neighbour_lists(ctx: *mut md_ctx, offsets: *const i64, partners: *mut i64) -> c_int;
    pub fn md_get_stats(ctx: *mut md_ctx, out: *mut md_stats) -> c_int;
    pub fn md_stream(ctx: *mut md_ctx) -> *mut c_void;
    pub fn md_synchronize(ctx: *mut md_ctx) -> c_int;
    pub fn md_time_kernels(ctx: *mut md_ctx, n_steps: i64, dt: f64, thermostat: *mut md_thermostat, barostat: *mut md_barostat, ms: *mut f64, launches: *mut i64) -> c_int;
    pub fn md_measure_fp64_peak(ctx: *mut md_ctx, tflops: *mut f64) -> c_int;
    pub fn md_invalidate_lists(ctx: *mut md_ctx) -> c_int;
}

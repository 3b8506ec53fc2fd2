// md_common.cuh — structs shared by host and device (Arrays, Params, Scalars, Mail/Peers), control-word helpers, pair geometry.
// Part of md_kernels.cuh (included from there, in order; one translation unit).
#pragma once

namespace md {


constexpr double K_B = 1.380648528;  // core/src/lib.rs:15

// Structure-of-arrays planes of the resident State (core/src/particle.rs:6-23), in cell-sorted order.
struct Arrays {
    double *x, *y, *z;     // Particle.position
    double *vx, *vy, *vz;  // Particle.velocity
    double *fx, *fy, *fz;  // Particle.force
    double *u;             // Particle.potential
    double *w;             // Particle.temp (Σ F_ij·r_ij)
    int *id;               // index of the particle in upload order
    double4 *q4;           // (x, y, z, -) packed copy for gathers in dense systems: one 32 B sector per partner
};

// Written by the host once per md_step / md_update_force call.
struct Params {
    double dt;         // delta_time
    double half_dt_m;  // delta_time / (2.0 * mass)        integrator.rs:30
    double mass;
    double sigma, eps, r_cut, u_cut;  // Potential::LennardJones  potential.rs:13-18
    double r_list;                    // r_cut + skin
    double th_tau, th_target;         // Thermostat::Berendsen{tau} + target temperature
    double ba_beta, ba_tau, ba_target;
    long long n;
    int th_kind, ba_kind;
};

// K5 slots: Σ m v (3), Σ m|v-c|², Σ m v·v, Σ W, Σ U, then the same COM/thermal sums for u = v + F c (the velocity
// right after the NEXT step's first half-kick: Nose-Hoover's second psi update needs its temperature,
// thermostat.rs:47-65), and last max |u|² (the displacement bound).  The max slot must stay last.
constexpr int NSUM = 12;
constexpr int S_MV = 0, S_TH = 3, S_KE = 4, S_W = 5, S_U = 6, S_MU = 7, S_THU = 10, S_MAX = NSUM - 1;

// Device-resident step state: box, thermostat/barostat coefficients, reduction results, loop control.
struct Scalars {
    double box[3];      // State::boundary_box
    double mu_pending;  // barostat.update's `position *= myu` not yet applied to x (1.0 = none)
    double lambda;      // Berendsen lambda for the step about to run (1.0 without thermostat)
    double mu;          // Berendsen myu for the step about to run (1.0 without barostat)
    double inv_scale;   // Π 1/myu since the last list build
    double disp_acc;    // upper bound of any atom's displacement since the list build (build-time units)
    double disp_next;   // upper bound of the next drift's displacement
    double shift[3];    // predicted COM velocity: shift of the one-pass thermal sum
    // last reduction
    double sum_mv[3], sum_th, sum_ke, sum_w, sum_u, max_w2;
    double vcom[3], thermal, kinetic, potential, temperature, pressure;
    double lambda_last, mu_last;  // coefficients used by the last executed step
    double psi;                   // Nose-Hoover friction after the last executed step (thermostat.rs:10-14)
    double temperature_mid;       // temperature of u = v + F c (after the next first half-kick, before scaling)
    long long steps_left, steps_done;
    int need_rebuild;
    int error;
    unsigned int ticket;
    int nbr_max;       // largest neighbour count of the last build
    int nbr_overflow;  // some atom exceeded the capacity
    int vel_is_half;   // 1: the velocity planes hold u = v + F*c (next step's first half-kick already applied)
    int out_of_box;    // the last cell binning saw a coordinate outside [0, L): list builds use the generic minimum image
    int tile_shell_max;  // tile kernels: largest shell (slots) and largest brick (atoms) of the last k_tile_measure
    int tile_own_max;
    int pad0;
    unsigned int bar_arrive;        // persistent step loop: arrivals at the mid-step grid barrier (only ever grows)
    unsigned int face_arrive[2];    // persistent step loop, multi-GPU: face blocks that have pushed their ghosts (only grows)
    unsigned long long epoch;  // multi-GPU peer-memory path: sequence number of the last finalized collective reduction
    unsigned long long wait_halo_ns, wait_sums_ns;  // time spent polling the mailboxes (block 0 / last block), accumulated
    unsigned long long t_start;                     // %globaltimer when the first block of the running k_force started
    unsigned long long force_atoms_ns, force_tail_ns;  // accumulated phase times (multi-GPU diagnostics)
    unsigned long long nbr_total;
    unsigned long long fin_seq;     // number of last-block epilogues completed so far (release-stored at their very end):
                                    // the end-of-step barrier of the persistent step loop
    // persistent step loop: accumulated phase times of block 0 (ns): drift, mid-step barrier, forces, reduction + finalize
    unsigned long long loop_ns[4];
    unsigned long long loop_steps;  // steps executed by the persistent loop since the last upload
    unsigned long long trace[16];   // MD_LOOP_TRACE builds: %globaltimer stamps of the last step's phases (md_debug_trace)
    double rank_sums[NSUM];  // multi-GPU: this rank's K5 sums (input of the all-gather)
    // multi-GPU rebuild bookkeeping
    int n_stay, n_left, n_right, n_lost;
    int g_left, g_right, pad1, pad2;
};

// Speculatively enqueued steps (multi-GPU chunks) turn into no-ops once the loop has to stop: every kernel of such a
// step checks this before touching anything.
__device__ __forceinline__ bool halted(const Scalars *sc)
{
    return sc->need_rebuild != 0 || sc->error != 0 || sc->steps_left <= 0;
}

// gpu-scope acquire / release accesses of the step-control words (L2, never a stale L1 line)
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Release / acquire on the atomic itself: one L2 round trip orders the thread's earlier writes (and, cumulatively, the
// writes of the block's other threads that a bar.sync placed before it) — where __threadfence() + atomicAdd() pays for a
// sequentially consistent fence (MEMBAR.SC.GPU) on top of the atomic.
__device__ __forceinline__ unsigned atom_add_release_gpu(unsigned *p, unsigned v)
{
    unsigned old;
    asm volatile("atom.add.release.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ unsigned atom_add_acq_rel_gpu(unsigned *p, unsigned v)
{
    unsigned old;
    asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(old) : "l"(p), "r"(v) : "memory");
    return old;
}
// halted(), read through L2: for a kernel that runs while its predecessor is still finishing
__device__ __forceinline__ bool halted_now(const Scalars *sc)
{
    return __ldcg(&sc->need_rebuild) != 0 || __ldcg(&sc->error) != 0 || __ldcg(&sc->steps_left) <= 0;
}

__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#ifdef MD_LOOP_TRACE
#define MD_TRACE(cond, k) do { if (cond) sc->trace[k] = gtime(); } while (0)
#else
#define MD_TRACE(cond, k) do { } while (0)
#endif
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// Per-thread asynchronous copies global → shared (LDGSTS): a thread parks the NEXT tile's operands in shared memory while
// it works on the current one, and reads back only what it copied itself — no barrier, no cross-thread hazard.
__device__ __forceinline__ void cp_async16(void *smem, const void *g)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async8(void *smem, const void *g)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem)), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// ---- multi-GPU peer-memory mailboxes (NVLink/NVSwitch, one process per GPU, buffers shared through CUDA IPC) ----------
// Every rank owns one Mail in its own HBM; the OTHER ranks write into it with plain stores over NVLink and the owner polls
// it locally.  Sequence numbers only grow, so nothing is ever reset; the sums are double-buffered by sequence parity because
// a rank may publish reduction s+1 while a non-neighbour is still folding reduction s.
constexpr int MAX_PEERS = 8;
struct Mail {
    // The rank sums travel as 8-byte words that carry their own flag (the low-latency protocol of collective libraries): a
    // word is {32 bits of payload, the low 32 bits of the reduction's sequence number}; an aligned 8-byte store arrives
    // whole, so the receiver polls the word itself and neither side needs a fence or a separate flag.  A double is two words.
    unsigned long long ll[2][MAX_PEERS][2 * NSUM];  // [seq & 1][source rank][2 * K5 slot + half]
    unsigned long long halo_seq[2];  // [0] left neighbour's, [1] right neighbour's drifted positions of step s are readable
};
struct Peers {      // lives in device memory; kernels get a pointer (NULL on one GPU)
    Mail *mail[MAX_PEERS];  // rank r's Mail as mapped into this process (mail[rank] is the local one)
    int rank, nranks;
    int left, right;        // ring neighbours (slab decomposition along x)
};

__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// Polls a local flag a peer writes.  Gives up after ~20 s (a peer died or the ranks diverged) so a broken run ends with an
// error instead of hanging the GPU.
__device__ __forceinline__ bool wait_seq(const unsigned long long *flag, unsigned long long seq)
{
    if (ld_acquire_sys(flag) >= seq) return true;
    const unsigned long long t0 = gtime();
    for (;;) {
        for (int spin = 0; spin < 64; ++spin)
            if (ld_acquire_sys(flag) >= seq) return true;
        if (gtime() - t0 > 20000000000ull) return false;
    }
}

struct Grid {
    int nc[3];
    int nsub;   // stencil half-width in cells
    int ncell;  // cells of the table (brick order pads the (x, y) columns of partial bricks: they stay empty)
    int cap;    // neighbour slots per atom
    int npad;   // row stride of the neighbour table
    // Brick order (dense systems, tile kernels — md_tile.cuh): the (x, y) columns of cells are numbered brick by brick
    // (4 x 4 columns each), z stays the fastest index.  An (x, y) brick cut into chunks of `bz` z-cells is the unit of work
    // of a thread block: its atoms and the +-2-cell shell around them are few contiguous runs of the sorted order.
    int brick;      // 0: column index cx * ncy + cy (canonical), 1: brick-major columns
    int nbx, nby;   // bricks along x and y
    int bz, nbz;    // z cells per chunk, chunks along z
};

__host__ __device__ __forceinline__ int col_index(const Grid &g, int cx, int cy)
{
    if (!g.brick) return cx * g.nc[1] + cy;
    return ((((cx >> 2) * g.nby) + (cy >> 2)) << 4) + ((cx & 3) << 2) + (cy & 3);
}

__host__ __device__ __forceinline__ int cell_index(const Grid &g, int cx, int cy, int cz)
{
    return col_index(g, cx, cy) * g.nc[2] + cz;
}

__host__ __device__ __forceinline__ void cell_decode(const Grid &g, int c, int &cx, int &cy, int &cz)
{
    cz = c % g.nc[2];
    const int col = c / g.nc[2];
    if (!g.brick) {
        cy = col % g.nc[1];
        cx = col / g.nc[1];
    } else {
        const int b = col >> 4, l = col & 15;
        cx = (b / g.nby) * 4 + (l >> 2);
        cy = (b % g.nby) * 4 + (l & 3);
    }
}


// ----------------------------------------------------------------------------------------------------
// Pair geometry shared by list build and force kernels: r = p_j - p_i with the reference's single-shift
// minimum image (potential.rs:181-200).  The comparisons are exact; only add/sub touch the FP64 pipe.
__device__ __forceinline__ double min_image(double r, double L, double h)
{
    if (r < -h) r = __dadd_rn(r, L);
    else if (r > h) r = __dsub_rn(r, L);
    return r;
}

// nalgebra Vector3::norm(): sqrt((x*x + y*y) + z*z), no contraction.
__device__ __forceinline__ double norm_exact(double rx, double ry, double rz)
{
    return __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz)));
}

}  // namespace md

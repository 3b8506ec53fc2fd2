// md_kernels.cuh — sm_100a kernels of the moldyn `solve` step loop (f64 throughout).
//
//   K1  k_cell_count / k_scan_* / k_scatter / k_sort_cells / k_reorder   cell binning + counting sort
//   K2  k_build_list                                                     Verlet-skin neighbour list
//   K3  k_force  (+ fused second half-kick and K5 partial sums)          potential.rs:158-216, integrator.rs:47-53
//   K4  k_kick_drift                                                     integrator.rs:28-45, barostat.rs:45-48
//   K5  block→warp deterministic reductions + finalize                   macro_parameters/*.rs, thermostat.rs, barostat.rs
//
// No tensor cores: the pair force is an irregular gather, not a dense contraction.  The streaming kernels are
// HBM-bound (coalesced SoA planes, 128-bit accesses), the force kernel is FP64-pipe / L1-gather bound.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
    int parity;        // fused one-kernel steps ping-pong x and v between two plane sets: which set is current
    int union_max;     // largest union-list length of the last k_build_union (entries per atom pair)
    int union_fail;    // k_build_union could not run (a coordinate outside the box): fall back to per-atom lists
    unsigned long long epoch;  // multi-GPU peer-memory path: sequence number of the last finalized collective reduction
    unsigned long long wait_halo_ns, wait_sums_ns;  // time spent polling the mailboxes (block 0 / last block), accumulated
    unsigned long long t_start;                     // %globaltimer when the first block of the running k_force started
    unsigned long long force_atoms_ns, force_tail_ns, drift_push_ns;  // accumulated phase times (multi-GPU diagnostics)
    unsigned long long nbr_total;
    unsigned long long probe[8];  // MD_TIMING_PROBES: %globaltimer stamps of k_force phases
    unsigned long long fin_seq;     // number of last-block epilogues completed so far (release-stored at their very end)
    unsigned long long chunk_fin0;  // fin_seq when the running step chunk started (early-start k_kick_drift, see there)
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

// Programmatic dependent launch (opt-in, MOLDYN_B200_PDL=1; single-GPU chunk graphs): a step kernel launched with the
// programmatic-serialization attribute becomes resident while its predecessor drains and blocks here until the predecessor
// has completed and its memory operations are visible.  Without the attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

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

#ifdef MD_TIMING_PROBES
#define PROBE(k) sc->probe[k] = gtime()
#define PROBE_MIN(k) atomicMin(&sc->probe[k], gtime())
#define PROBE_MAX(k) atomicMax(&sc->probe[k], gtime())
#else
#define PROBE(k)
#define PROBE_MIN(k)
#define PROBE_MAX(k)
#endif

// ---- multi-GPU peer-memory mailboxes (NVLink/NVSwitch, one process per GPU, buffers shared through CUDA IPC) ----------
// Every rank owns one Mail in its own HBM; the OTHER ranks write into it with plain stores over NVLink and the owner polls
// it locally.  Sequence numbers only grow, so nothing is ever reset; the sums are double-buffered by sequence parity because
// a rank may publish reduction s+1 while a non-neighbour is still folding reduction s.
constexpr int MAX_PEERS = 8;
struct Mail {
    double sums[2][MAX_PEERS][12];          // [seq & 1][source rank][K5 slot]
    unsigned long long sums_seq[MAX_PEERS];  // sums_seq[r] = s: rank r's sums of reduction s have landed
    unsigned long long halo_seq[2];          // [0] left neighbour's, [1] right neighbour's ghost positions of step s landed
};
static_assert(NSUM == 12, "Mail::sums holds NSUM slots per rank");
struct Peers {      // lives in device memory; kernels get a pointer (NULL on one GPU)
    Mail *mail[MAX_PEERS];  // rank r's Mail as mapped into this process (mail[rank] is the local one)
    int rank, nranks;
    int left, right;        // ring neighbours (slab decomposition along x)
};

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
    int ncell;
    int cap;    // neighbour slots per atom
    int npad;   // row stride of the neighbour table
};


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

// ----------------------------------------------------------------------------------------------------
// K1: cell index.  c_d = min(nc_d - 1, (int)(frac(x_d / L_d) * nc_d)); positions outside the box are
// binned by their periodic image (the force arithmetic itself never wraps them — the reference does not).
__device__ __forceinline__ int cell_coord(double x, double L, int nc)
{
    double s = __ddiv_rn(x, L);
    s = __dsub_rn(s, floor(s));
    int c = (int)__dmul_rn(s, (double)nc);
    return min(max(c, 0), nc - 1);
}

__global__ void k_cell_count(int n, const double *__restrict__ x, const double *__restrict__ y,
                             const double *__restrict__ z, Scalars *sc, Grid g,
                             int *__restrict__ cell_of, int *__restrict__ cell_cnt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int cx = cell_coord(x[i], sc->box[0], g.nc[0]);
    int cy = cell_coord(y[i], sc->box[1], g.nc[1]);
    int cz = cell_coord(z[i], sc->box[2], g.nc[2]);
    int c = (cx * g.nc[1] + cy) * g.nc[2] + cz;
    cell_of[i] = c;
    atomicAdd(&cell_cnt[c], 1);
    // (inside md_step the drift kernel keeps every coordinate in [0, L); an uploaded State may hold anything)
    const double xx = x[i], yy = y[i], zz = z[i];
    if (xx < 0.0 || xx >= sc->box[0] || yy < 0.0 || yy >= sc->box[1] || zz < 0.0 || zz >= sc->box[2] || xx != xx ||
        yy != yy || zz != zz)
        sc->out_of_box = 1;
}

// Exclusive scan of cell counts: per-block scan (1024 items) → scan of block totals → add back.
constexpr int SCAN_BLOCK = 1024;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total)
{
    __shared__ int warp_sums[32];
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int ws = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += t;
        }
        warp_sums[lane] = ws;
    }
    __syncthreads();
    int base = wid ? warp_sums[wid - 1] : 0;
    *total = warp_sums[31];
    __syncthreads();
    return base + inc - v;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_block(int n, const int *__restrict__ in,
                                                           int *__restrict__ out, int *__restrict__ block_sums)
{
    int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    int v = (i < n) ? in[i] : 0;
    int total;
    int ex = block_exclusive_scan(v, &total);
    if (i < n) out[i] = ex;
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_sums(int nblocks, int *__restrict__ block_sums)
{
    __shared__ int carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < nblocks; base += SCAN_BLOCK) {
        int i = base + threadIdx.x;
        int v = (i < nblocks) ? block_sums[i] : 0;
        int total;
        int ex = block_exclusive_scan(v, &total);
        int carry = carry_s;
        if (i < nblocks) block_sums[i] = ex + carry;
        __syncthreads();
        if (threadIdx.x == 0) carry_s = carry + total;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_BLOCK) k_scan_add(int n, int *__restrict__ out,
                                                         const int *__restrict__ block_sums, int total_items)
{
    int i = blockIdx.x * SCAN_BLOCK + threadIdx.x;
    if (i < n) out[i] += block_sums[blockIdx.x];
    if (i == 0 && total_items >= 0) out[n] = total_items;
}

// total of an exclusive scan whose item count is not known to the host: out[n] = out[n-1] + in[n-1]
__global__ void k_scan_total(int n, const int *__restrict__ in, int *__restrict__ out)
{
    out[n] = n > 0 ? out[n - 1] + in[n - 1] : 0;
}

__global__ void k_scatter(int n, const int *__restrict__ cell_of, const int *__restrict__ cell_start,
                          int *__restrict__ cell_fill, int *__restrict__ order)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int c = cell_of[i];
    int slot = cell_start[c] + atomicAdd(&cell_fill[c], 1);
    order[slot] = i;
}

// Makes the order inside every cell independent of atomic arrival order: ascending upload index.
// The sorted order of the whole system is then the lexicographic (cell, upload index) order — deterministic.
__global__ void k_sort_cells(int ncell, const int *__restrict__ cell_start, const int *__restrict__ id_old,
                             int *__restrict__ order)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    int s = cell_start[c], e = cell_start[c + 1];
    for (int a = s + 1; a < e; ++a) {
        int item = order[a];
        int key = id_old[item];
        int b = a - 1;
        while (b >= s && id_old[order[b]] > key) {
            order[b + 1] = order[b];
            --b;
        }
        order[b + 1] = item;
    }
}

__global__ void k_reorder(int n, const int *__restrict__ order, const int *__restrict__ cell_of, Arrays src,
                          Arrays dst, int *__restrict__ cell_sorted)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    int s = order[p];
    dst.x[p] = src.x[s];   dst.y[p] = src.y[s];   dst.z[p] = src.z[s];
    dst.vx[p] = src.vx[s]; dst.vy[p] = src.vy[s]; dst.vz[p] = src.vz[s];
    dst.fx[p] = src.fx[s]; dst.fy[p] = src.fy[s]; dst.fz[p] = src.fz[s];
    dst.u[p] = src.u[s];   dst.w[p] = src.w[s];
    dst.id[p] = src.id[s];
    cell_sorted[p] = cell_of[s];
}

// ----------------------------------------------------------------------------------------------------
// K2: Verlet list.  One thread per atom walks the (deduplicated) cell stencil and keeps partners whose
// reference min-image distance is <= r_list — the predicate of potential.rs:181-204 widened by the skin,
// evaluated in the reference's exact arithmetic so the pair set is the reference's, bit for bit.
// Table layout nbr[k * npad + p]: a warp reads one coalesced row per k.
// SHIFT: every dimension has at least 2*nsub + 3 cells.  Then a stencil cell that was wrapped around the box holds exactly
// the partners the reference's single-shift rule moves by -/+L (|x_q - x_i| > L/2 there and < L/2 everywhere else), so the
// image shift is a constant of the cell run — added with the reference's own operation (r + L, r - L; adding 0.0 is exact) —
// and the two compares per axis and candidate of min_image() disappear from the inner loop (~100 → ~30 instructions).
template <bool SORT_BY_ID, bool SHIFT>
__global__ void __launch_bounds__(128) k_build_list(int n, Arrays a, const int *__restrict__ cell_sorted,
                                                    const int *__restrict__ cell_start, Scalars *sc, Grid g,
                                                    double r_list, double r2_list, int *__restrict__ nbr,
                                                    int *__restrict__ nbr_cnt)
{
    // r2_list is the largest double whose correctly rounded square root is <= r_list (computed by the host), so
    // `(rx*rx + ry*ry) + rz*rz <= r2_list` IS the predicate `norm(r) <= r_list` — without the square root.
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    if (p < n) {
        const double Lx = sc->box[0], Ly = sc->box[1], Lz = sc->box[2];
        const double hx = Lx / 2.0, hy = Ly / 2.0, hz = Lz / 2.0;
        const double xi = a.x[p], yi = a.y[p], zi = a.z[p];
        const int ncx = g.nc[0], ncy = g.nc[1], ncz = g.nc[2];
        int c = cell_sorted[p];
        int cz = c % ncz;
        int cy = (c / ncz) % ncy;
        int cx = c / (ncz * ncy);
        int w = 2 * g.nsub + 1;
        int lox, loy, nx, ny;
        if (ncx >= w) { lox = cx - g.nsub; nx = w; } else { lox = 0; nx = ncx; }
        if (ncy >= w) { loy = cy - g.nsub; ny = w; } else { loy = 0; ny = ncy; }
        // z is the fastest cell index, so the z-stencil of one (x,y) column is at most two contiguous runs of cells
        int z0a, z1a, z0b = 0, z1b = 0;  // half-open cell ranges [z0, z1)
        if (ncz >= w) {
            int lo = cz - g.nsub, hi = cz + g.nsub + 1;
            if (lo < 0) { z0a = 0; z1a = hi; z0b = lo + ncz; z1b = ncz; }
            else if (hi > ncz) { z0a = lo; z1a = ncz; z0b = 0; z1b = hi - ncz; }
            else { z0a = lo; z1a = hi; }
        } else { z0a = 0; z1a = ncz; }
        const bool in_box = sc->out_of_box == 0;
        // which z run holds the wrapped cells, and which way the reference's rule shifts their atoms
        const double szb = (ncz >= w && cz - g.nsub < 0) ? -Lz : Lz;
        for (int ia = 0; ia < nx; ++ia) {
            int qx = lox + ia;
            const double sx = qx < 0 ? -Lx : (qx >= ncx ? Lx : 0.0);
            qx += (qx < 0) ? ncx : 0;
            qx -= (qx >= ncx) ? ncx : 0;
            for (int ib = 0; ib < ny; ++ib) {
                int qy = loy + ib;
                const double sy = qy < 0 ? -Ly : (qy >= ncy ? Ly : 0.0);
                qy += (qy < 0) ? ncy : 0;
                qy -= (qy >= ncy) ? ncy : 0;
                const int base = (qx * ncy + qy) * ncz;
                // both runs' bounds are fetched before either is walked
                const int sa = cell_start[base + z0a], ea = cell_start[base + z1a];
                const int sb = (z1b > z0b) ? cell_start[base + z0b] : 0, eb = (z1b > z0b) ? cell_start[base + z1b] : 0;
#pragma unroll 1
                for (int run = 0; run < 2; ++run) {
                    const int s = run ? sb : sa, e = run ? eb : ea;
                    if (SHIFT && in_box) {
                        const double sz = run ? szb : 0.0;
                        for (int q = s; q < e; ++q) {
                            const double rx = __dadd_rn(__dsub_rn(a.x[q], xi), sx);
                            if (fabs(rx) > r_list) continue;
                            const double ry = __dadd_rn(__dsub_rn(a.y[q], yi), sy);
                            const double rz = __dadd_rn(__dsub_rn(a.z[q], zi), sz);
                            const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
                            if (r2 > r2_list || q == p) continue;
                            if (cnt < g.cap) nbr[(size_t)cnt * g.npad + p] = q;
                            ++cnt;
                        }
                    } else {
                        for (int q = s; q < e; ++q) {
                            double rx = min_image(__dsub_rn(a.x[q], xi), Lx, hx);
                            if (fabs(rx) > r_list) continue;
                            double ry = min_image(__dsub_rn(a.y[q], yi), Ly, hy);
                            double rz = min_image(__dsub_rn(a.z[q], zi), Lz, hz);
                            double r2 = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
                            if (r2 > r2_list || q == p) continue;
                            if (cnt < g.cap) nbr[(size_t)cnt * g.npad + p] = q;
                            ++cnt;
                        }
                    }
                }
            }
        }
        nbr_cnt[p] = min(cnt, g.cap);
        if (SORT_BY_ID && cnt <= g.cap) {
            // ascending upload index == the reference's ascending j (potential.rs:177)
            for (int s1 = 1; s1 < cnt; ++s1) {
                int item = nbr[(size_t)s1 * g.npad + p];
                int key = a.id[item];
                int b = s1 - 1;
                while (b >= 0 && a.id[nbr[(size_t)b * g.npad + p]] > key) {
                    nbr[(size_t)(b + 1) * g.npad + p] = nbr[(size_t)b * g.npad + p];
                    --b;
                }
                nbr[(size_t)(b + 1) * g.npad + p] = item;
            }
        }
    }
    // statistics: max / total / overflow (integer atomics — order-independent results)
    int wmax = cnt;
    unsigned int wsum = (unsigned int)cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    }
    if ((threadIdx.x & 31) == 0 && wsum) {
        atomicMax(&sc->nbr_max, wmax);
        atomicAdd(&sc->nbr_total, (unsigned long long)wsum);
        if (wmax > g.cap) atomicExch(&sc->nbr_overflow, 1);
    }
}

// ----------------------------------------------------------------------------------------------------
// K2 for dense systems, FAST mode: UNION lists.  The force kernel gives two consecutive (cell-sorted, hence spatially
// adjacent) atoms A = 2t, B = 2t+1 to one thread, and the dense loop is bound by the L1's gather rate (one pass per lane and
// partner).  A and B share ~3/4 of their partners, so thread t gets ONE list: every atom within r_list of A or of B, each
// entry tagged with two membership bits (bit 30: in A's list, bit 31: in B's).  A partner is then gathered once and
// evaluated against both atoms: ~1.25x the pair arithmetic for ~0.63x the gathers.  The membership bits make the union
// exactly equivalent to the two per-atom lists (md_neighbour_lists reconstructs them from the bits).
//
// One thread scans A's stencil once, testing both atoms, then the cells of B's stencil that A's stencil does not cover
// (B only).  Requires >= 2*nsub + 5 cells per dimension and every coordinate inside the box: the periodic image of a stencil
// cell is then a per-run constant for A (cell wrap) and for B (nearest image by cell distance), added with the reference's
// own r + L / r - L — same exact predicate as k_build_list<.., SHIFT = true>.
constexpr int UNION_A = 1 << 30;
constexpr unsigned int UNION_B = 1u << 31;
constexpr int UNION_IDX = (1 << 30) - 1;

__device__ __forceinline__ double image_shift(int q, int b, int nc, double L)
{
    const int d = q - b;
    return 2 * d > nc ? -L : (2 * d < -nc ? L : 0.0);
}

__device__ __forceinline__ int cell_dist(int q, int a, int nc)
{
    const int d = abs(q - a);
    return min(d, nc - d);
}

__global__ void __launch_bounds__(128) k_build_union(int n, Arrays a, const int *__restrict__ cell_sorted,
                                                     const int *__restrict__ cell_start, Scalars *sc, Grid g,
                                                     double r_list, double r2_list, int *__restrict__ nbr_u,
                                                     int *__restrict__ cnt_u, int cap_u, int pstride,
                                                     int *__restrict__ nbr_cnt)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int npairs = (n + 1) >> 1;
    int cnt = 0, cnt_a = 0, cnt_b = 0;
    if (sc->out_of_box) {
        if (t == 0) sc->union_fail = 1;
        return;
    }
    if (t < npairs) {
        const int A = 2 * t;
        const bool has_b = A + 1 < n;
        const int B = has_b ? A + 1 : A;
        const double Lx = sc->box[0], Ly = sc->box[1], Lz = sc->box[2];
        const double xa = a.x[A], ya = a.y[A], za = a.z[A];
        const double xb = a.x[B], yb = a.y[B], zb = a.z[B];
        const int ncx = g.nc[0], ncy = g.nc[1], ncz = g.nc[2], ns = g.nsub, w = 2 * g.nsub + 1;
        const int ca = cell_sorted[A], cb = cell_sorted[B];
        const int az = ca % ncz, ay = (ca / ncz) % ncy, ax = ca / (ncz * ncy);
        const int bz = cb % ncz, by = (cb / ncz) % ncy, bx = cb / (ncz * ncy);
        auto emit = [&](int q, bool in_a, bool in_b) {
            if (cnt < cap_u) nbr_u[(size_t)cnt * pstride + t] = (int)((unsigned int)q | (in_a ? (unsigned int)UNION_A : 0u) | (in_b ? UNION_B : 0u));
            ++cnt;
            cnt_a += in_a ? 1 : 0;
            cnt_b += in_b ? 1 : 0;
        };
        // ---- pass 1: A's stencil, both atoms ----
        int z0a, z1a, z0b = 0, z1b = 0;
        {
            const int lo = az - ns, hi = az + ns + 1;
            if (lo < 0) { z0a = 0; z1a = hi; z0b = lo + ncz; z1b = ncz; }
            else if (hi > ncz) { z0a = lo; z1a = ncz; z0b = 0; z1b = hi - ncz; }
            else { z0a = lo; z1a = hi; }
        }
        const double sza_b = (az - ns < 0) ? -Lz : Lz;  // A's shift for the wrapped z run
        for (int ia = 0; ia < w; ++ia) {
            int qx = ax - ns + ia;
            const double sxa = qx < 0 ? -Lx : (qx >= ncx ? Lx : 0.0);
            qx += (qx < 0) ? ncx : 0;
            qx -= (qx >= ncx) ? ncx : 0;
            const double sxb = image_shift(qx, bx, ncx, Lx);
            for (int ib = 0; ib < w; ++ib) {
                int qy = ay - ns + ib;
                const double sya = qy < 0 ? -Ly : (qy >= ncy ? Ly : 0.0);
                qy += (qy < 0) ? ncy : 0;
                qy -= (qy >= ncy) ? ncy : 0;
                const double syb = image_shift(qy, by, ncy, Ly);
                const int base = (qx * ncy + qy) * ncz;
                const int sa = cell_start[base + z0a], ea = cell_start[base + z1a];
                const int sb = (z1b > z0b) ? cell_start[base + z0b] : 0, eb = (z1b > z0b) ? cell_start[base + z1b] : 0;
#pragma unroll 1
                for (int run = 0; run < 2; ++run) {
                    const int s = run ? sb : sa, e = run ? eb : ea;
                    const double sza = run ? sza_b : 0.0;
                    const double szb = image_shift(run ? z0b : z0a, bz, ncz, Lz);  // constant over a run (>= 2ns+5 cells)
                    for (int q = s; q < e; ++q) {
                        const double xq = a.x[q];
                        const double rxa = __dadd_rn(__dsub_rn(xq, xa), sxa), rxb = __dadd_rn(__dsub_rn(xq, xb), sxb);
                        if (fabs(rxa) > r_list && fabs(rxb) > r_list) continue;
                        const double yq = a.y[q], zq = a.z[q];
                        const double rya = __dadd_rn(__dsub_rn(yq, ya), sya), rza = __dadd_rn(__dsub_rn(zq, za), sza);
                        const double ryb = __dadd_rn(__dsub_rn(yq, yb), syb), rzb = __dadd_rn(__dsub_rn(zq, zb), szb);
                        const double r2a = __dadd_rn(__dadd_rn(__dmul_rn(rxa, rxa), __dmul_rn(rya, rya)), __dmul_rn(rza, rza));
                        const double r2b = __dadd_rn(__dadd_rn(__dmul_rn(rxb, rxb), __dmul_rn(ryb, ryb)), __dmul_rn(rzb, rzb));
                        const bool in_a = r2a <= r2_list && q != A;
                        const bool in_b = has_b && r2b <= r2_list && q != B;
                        if (in_a || in_b) emit(q, in_a, in_b);
                    }
                }
            }
        }
        // ---- pass 2: cells of B's stencil outside A's stencil, B only ----
        if (has_b && cb != ca) {
            for (int ia = 0; ia < w; ++ia) {
                int qx = bx - ns + ia;
                const double sx = qx < 0 ? -Lx : (qx >= ncx ? Lx : 0.0);
                qx += (qx < 0) ? ncx : 0;
                qx -= (qx >= ncx) ? ncx : 0;
                const bool in_x = cell_dist(qx, ax, ncx) <= ns;
                for (int ib = 0; ib < w; ++ib) {
                    int qy = by - ns + ib;
                    const double sy = qy < 0 ? -Ly : (qy >= ncy ? Ly : 0.0);
                    qy += (qy < 0) ? ncy : 0;
                    qy -= (qy >= ncy) ? ncy : 0;
                    const bool in_xy = in_x && cell_dist(qy, ay, ncy) <= ns;
                    for (int ic = 0; ic < w; ++ic) {
                        int qz = bz - ns + ic;
                        const double sz = qz < 0 ? -Lz : (qz >= ncz ? Lz : 0.0);
                        qz += (qz < 0) ? ncz : 0;
                        qz -= (qz >= ncz) ? ncz : 0;
                        if (in_xy && cell_dist(qz, az, ncz) <= ns) continue;  // pass 1 has seen this cell
                        const int c = (qx * ncy + qy) * ncz + qz;
                        for (int q = cell_start[c]; q < cell_start[c + 1]; ++q) {
                            const double rx = __dadd_rn(__dsub_rn(a.x[q], xb), sx);
                            if (fabs(rx) > r_list) continue;
                            const double ry = __dadd_rn(__dsub_rn(a.y[q], yb), sy), rz = __dadd_rn(__dsub_rn(a.z[q], zb), sz);
                            const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
                            if (r2 <= r2_list && q != B) emit(q, false, true);
                        }
                    }
                }
            }
        }
        cnt_u[t] = min(cnt, cap_u);
        nbr_cnt[A] = cnt_a;
        if (has_b) nbr_cnt[B] = cnt_b;
    }
    // statistics (integer atomics — order-independent): per-atom max / total as for k_build_list, plus the union length
    int wmax = max(cnt_a, cnt_b), umax = cnt;
    unsigned int wsum = (unsigned int)(cnt_a + cnt_b);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        umax = max(umax, __shfl_xor_sync(0xffffffffu, umax, o));
        wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    }
    if ((threadIdx.x & 31) == 0 && wsum) {
        atomicMax(&sc->nbr_max, wmax);
        atomicMax(&sc->union_max, umax);
        atomicAdd(&sc->nbr_total, (unsigned long long)wsum);
        if (umax > cap_u) atomicExch(&sc->nbr_overflow, 1);
    }
}

// ----------------------------------------------------------------------------------------------------
// K5: deterministic reductions.  Lane tree (xor shuffles) → fixed-order sum over warps → one slot per block;
// the last block to finish (atomic ticket) folds the per-block slots in a fixed order and finalizes.
struct Sums {
    double v[NSUM];
};

__device__ __forceinline__ void warp_reduce(Sums &s)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < NSUM - 1; ++q) s.v[q] += __shfl_xor_sync(0xffffffffu, s.v[q], o);
        s.v[NSUM - 1] = fmax(s.v[NSUM - 1], __shfl_xor_sync(0xffffffffu, s.v[NSUM - 1], o));
    }
}

// All threads of the block must call. Result valid in thread 0.
template <int BLOCK>
__device__ __forceinline__ void block_reduce(Sums &s)
{
    __shared__ double sm[BLOCK / 32][NSUM];
    warp_reduce(s);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NSUM; ++q) sm[wid][q] = s.v[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < BLOCK / 32; ++w) {
#pragma unroll
            for (int q = 0; q < NSUM - 1; ++q) s.v[q] += sm[w][q];
            s.v[NSUM - 1] = fmax(s.v[NSUM - 1], sm[w][NSUM - 1]);
        }
    }
    __syncthreads();
}

// Thermostat coefficient of the NEXT step.  Berendsen (thermostat.rs:31-34): lambda from the temperature at the step
// start.  Nose-Hoover (thermostat.rs:35-39, 59-65): psi advances by half a step with the start temperature, lambda =
// exp(-psi dt/2), then psi advances again with the temperature after the first half-kick (before scaling).
__device__ __forceinline__ double thermostat_lambda(int kind, double dt, double tau, double target, double t_start,
                                                    double t_mid, double &psi)
{
    if (kind == 1) return sqrt(1.0 + dt / tau * (target / t_start - 1.0));
    if (kind == 2) {
        double psi_dot = -((target / t_start) - 1.0) / tau;
        psi += psi_dot * (dt / 2.0);
        const double lambda = exp(-psi * dt / 2.0);
        psi_dot = -((target / t_mid) - 1.0) / tau;
        psi += psi_dot * (dt / 2.0);
        return lambda;
    }
    return 1.0;
}

// Step controls for the first step of a batch from the stored macro state (thermostat.rs:24-44, barostat.rs:21-31)
// plus the displacement bookkeeping that triggers list rebuilds.  psi_in: the caller's Nose-Hoover state.
__device__ __forceinline__ void compute_controls(Scalars *sc, const Params *pr, double psi_in)
{
    double mu = 1.0, psi = psi_in;
    const double lambda = thermostat_lambda(pr->th_kind, pr->dt, pr->th_tau, pr->th_target, sc->temperature,
                                            sc->temperature_mid, psi);
    if (pr->ba_kind == 1) {
        double myu_cubed = 1.0 + pr->dt * pr->ba_beta / pr->ba_tau * (sc->pressure - pr->ba_target);
        mu = cbrt(myu_cubed);
    }
    sc->lambda = lambda;
    sc->mu = mu;
    sc->psi = psi;
    // ΣF = 0, so the COM velocity after the next step's kicks is lambda * vcom: used as the shift that keeps
    // the one-pass thermal sum Σ m|v-c|² free of cancellation.
    sc->shift[0] = sc->vcom[0] * lambda;
    sc->shift[1] = sc->vcom[1] * lambda;
    sc->shift[2] = sc->vcom[2] * lambda;
    // Next drift moves every atom by at most lambda*sqrt(max|v + F c|²)*dt; in build-time units that is
    // multiplied by inv_scale (positions and box have been scaled by Π myu since the build).
    double vmax = lambda * sqrt(sc->max_w2);
    sc->disp_next = vmax * pr->dt * sc->inv_scale;
    // Pair now within r_cut ⇒ at build time within r_cut*inv_scale + 2*disp ≤ r_list must hold.
    double thr = 0.5 * (pr->r_list - pr->r_cut * sc->inv_scale) * (1.0 - 1e-9);
    double d = sc->disp_acc + sc->disp_next;
    sc->need_rebuild = (d > thr) ? 1 : 0;
    if (!(d == d) || !(lambda == lambda) || !(mu == mu) || isinf(d) || isinf(lambda) || isinf(mu)) sc->error = 7;
}

// mode bits of finalize
constexpr int FIN_STEP = 1;  // called at the end of an MD step: commit drift, apply barostat box scaling, count
constexpr int FIN_DIST = 2;  // multi-GPU: publish this rank's sums only; k_finalize_dist finalizes after the all-gather
constexpr int FIN_FLIP = 4;  // fused one-kernel step: the step wrote the other plane set, flip sc->parity
constexpr int FIN_P2P = 8;   // multi-GPU: exchange the rank sums through the peer mailboxes and finalize right here

// `in` / `pr`: the control words and parameters as they were when the kernel started (the last block copies them into
// shared memory while it waits for the partial sums, so finalize starts without a trip to global memory); `sc`: where the
// results go.  The two may alias (k_finalize_dist).
__device__ __forceinline__ void finalize(Scalars *sc, const Scalars *in, const Params *pr, const Sums &t, int mode)
{
    const double n = (double)pr->n, mass = pr->mass, dt = pr->dt, r_list = pr->r_list, r_cut = pr->r_cut;
    const int th_kind = pr->th_kind, ba_kind = pr->ba_kind;
    const double th_tau = pr->th_tau, th_target = pr->th_target;
    const double ba_beta = pr->ba_beta, ba_tau = pr->ba_tau, ba_target = pr->ba_target;
    double box0 = in->box[0], box1 = in->box[1], box2 = in->box[2];
    const double shift0 = in->shift[0], shift1 = in->shift[1], shift2 = in->shift[2];
    const double lambda_used = in->lambda, mu_used = in->mu;
    double disp_acc = in->disp_acc, inv_scale = in->inv_scale;
    const double disp_next_old = in->disp_next;
    const long long steps_left = in->steps_left, steps_done = in->steps_done;
    double psi = in->psi;

    const double M = n * mass;
    const double vc0 = t.v[0] / M, vc1 = t.v[1] / M, vc2 = t.v[2] / M;  // get_center_of_mass_velocity  mod.rs:12-25
    const double e0 = vc0 - shift0, e1 = vc1 - shift1, e2 = vc2 - shift2;
    const double th2 = t.v[3] - M * (e0 * e0 + e1 * e1 + e2 * e2);       // Σ m |v - vcom|²
    const double thermal = th2 / 2.0;                                    // get_thermal_energy   energy.rs:25-37
    if (mode & FIN_STEP) {
        disp_acc += disp_next_old;  // the drift that preceded this force evaluation
        if (ba_kind == 1) {         // barostat.update: boundary_box *= myu  (barostat.rs:45); x *= myu is deferred
            box0 *= mu_used; box1 *= mu_used; box2 *= mu_used;
            inv_scale /= mu_used;
        }
    }
    const double temperature = (2.0 * thermal) / (3.0 * n * K_B) * 100.0;  // temperature.rs:4-7
    // same for u = v + F c (the state thermostat.update sees after the next first half-kick)
    // (the u sums are only accumulated when something reads them: Nose-Hoover, or a plain force evaluation)
    double temperature_mid = temperature;
    if (th_kind == 2 || !(mode & FIN_STEP)) {
        const double uc0 = t.v[S_MU] / M - shift0, uc1 = t.v[S_MU + 1] / M - shift1, uc2 = t.v[S_MU + 2] / M - shift2;
        const double thu2 = t.v[S_THU] - M * (uc0 * uc0 + uc1 * uc1 + uc2 * uc2);
        temperature_mid = (2.0 * (thu2 / 2.0)) / (3.0 * n * K_B) * 100.0;
    }
    const double volume = box0 * box1 * box2;
    const double pressure = (th2 + (-t.v[5]) * 0.5) / volume / 3.0;         // pressure.rs:5-20
    // controls of the NEXT step (thermostat.rs:24-34, barostat.rs:21-31)
    // (Nose-Hoover's psi is only advanced when this batch has a next step: the first step of the next batch is
    // prepared by k_prepare from the caller's psi.)
    double lambda = 1.0, mu = 1.0;
    const bool more = !(mode & FIN_STEP) || steps_left - 1 > 0;
    if (th_kind == 1 || (th_kind == 2 && more))
        lambda = thermostat_lambda(th_kind, dt, th_tau, th_target, temperature, temperature_mid, psi);
    if (ba_kind == 1) mu = cbrt(1.0 + dt * ba_beta / ba_tau * (pressure - ba_target));
    const double vmax = lambda * sqrt(t.v[S_MAX]);
    const double disp_next = vmax * dt * inv_scale;
    const double thr = 0.5 * (r_list - r_cut * inv_scale) * (1.0 - 1e-9);
    const double d = disp_acc + disp_next;

    for (int k = 0; k < 3; ++k) sc->sum_mv[k] = t.v[k];
    sc->sum_th = t.v[3]; sc->sum_ke = t.v[4]; sc->sum_w = t.v[5]; sc->sum_u = t.v[6]; sc->max_w2 = t.v[S_MAX];
    sc->vcom[0] = vc0; sc->vcom[1] = vc1; sc->vcom[2] = vc2;
    sc->thermal = thermal;
    sc->kinetic = t.v[4] / 2.0;    // get_kinetic_energy   energy.rs:14-22
    sc->potential = t.v[6] / 2.0;  // get_potential_energy energy.rs:40-49
    sc->temperature = temperature;
    sc->temperature_mid = temperature_mid;
    sc->pressure = pressure;
    if (mode & FIN_STEP) sc->psi = psi;
    if (mode & FIN_STEP) {
        sc->lambda_last = lambda_used;
        sc->mu_last = mu_used;
        if (ba_kind == 1) {
            sc->box[0] = box0; sc->box[1] = box1; sc->box[2] = box2;
            sc->mu_pending = mu_used;
        }
        sc->steps_left = steps_left - 1;
        sc->steps_done = steps_done + 1;
        // k_force stored u = v + F*c instead of v unless this was the last step of the batch
        sc->vel_is_half = steps_left - 1 > 0 ? 1 : 0;
        if (mode & FIN_FLIP) sc->parity ^= 1;
    }
    sc->disp_acc = disp_acc;
    sc->inv_scale = inv_scale;
    sc->lambda = lambda;
    sc->mu = mu;
    // ΣF = 0, so the COM velocity after the next step's kicks is lambda * vcom: the shift that keeps the one-pass
    // thermal sum Σ m|v-c|² free of cancellation.
    sc->shift[0] = vc0 * lambda; sc->shift[1] = vc1 * lambda; sc->shift[2] = vc2 * lambda;
    // Next drift moves every atom by at most lambda*sqrt(max|v + F c|²)*dt (in build-time units: x inv_scale).
    // A pair now within r_cut must have been within r_cut*inv_scale + 2*disp <= r_list at build time.
    sc->disp_next = disp_next;
    sc->need_rebuild = (d > thr) ? 1 : 0;
    if (!(d == d) || !(lambda == lambda) || !(mu == mu) || isinf(d) || isinf(lambda) || isinf(mu)) sc->error = 7;
}

// Last-block epilogue shared by k_force and k_reduce_state.  `mine` is this block's reduced sums (thread 0).
template <int BLOCK>
__device__ __forceinline__ void grid_reduce_finalize(Sums &mine, double *__restrict__ partials, Scalars *sc,
                                                     const Params *pr, int mode,
                                                     unsigned long long cond_handle, const Peers *peers_p)
{
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NSUM; ++q) __stcg(&partials[(size_t)blockIdx.x * NSUM + q], mine.v[q]);
        __threadfence();
        unsigned int t = atomicAdd(&sc->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    if (threadIdx.x == 0) { PROBE(2); }
    __threadfence();
    // Last block.  (1) A copy of the control words and parameters finalize reads goes to shared memory — those loads are in
    // flight together with (2) the fold of the per-block partials: thread (g, q) adds slot q of blocks g, g+G, g+2G, … in
    // ascending order (independent loads, one L2 round trip), then the G group sums of a slot are added in group order.
    // Fixed assignment, fixed order: the result depends on the grid size only.
    constexpr int H = NSUM / 2;   // slot pairs: 128-bit loads
    constexpr int G = BLOCK / H;  // groups of blocks
    static_assert(NSUM % 2 == 0, "slots are folded in pairs");
    __shared__ double fold[G][NSUM];
    __shared__ double folded[NSUM];
    __shared__ Scalars sc_in;
    __shared__ Params pr_in;
    {
        constexpr int WS = (int)(sizeof(Scalars) / 8), WP = (int)(sizeof(Params) / 8);
        static_assert(sizeof(Scalars) % 8 == 0 && sizeof(Params) % 8 == 0, "copied as 64-bit words");
        const unsigned long long *gs = reinterpret_cast<const unsigned long long *>(sc);
        const unsigned long long *gp = reinterpret_cast<const unsigned long long *>(pr);
        unsigned long long *ss_ = reinterpret_cast<unsigned long long *>(&sc_in), *sp_ = reinterpret_cast<unsigned long long *>(&pr_in);
        for (int w = threadIdx.x; w < WS + WP; w += BLOCK) {
            if (w < WS) ss_[w] = __ldcg(gs + w);
            else sp_[w - WS] = __ldcg(gp + (w - WS));
        }
    }
    {
        const int h = threadIdx.x % H, g = threadIdx.x / H;
        if (g < G) {
            const bool has_max = (h == H - 1);  // the last slot of the last pair is the running maximum
            double ax = 0.0, ay = 0.0;
            const double2 *src = reinterpret_cast<const double2 *>(partials) + h;
            constexpr int U = 16;               // loads in flight per thread
            unsigned int b = g;
            for (; b + (U - 1) * G < gridDim.x; b += U * G) {
                double2 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = __ldcg(src + (size_t)(b + u * G) * H);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    ax += v[u].x;
                    ay = has_max ? fmax(ay, v[u].y) : ay + v[u].y;
                }
            }
            {
                double2 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned int bb = b + u * G;
                    v[u] = bb < gridDim.x ? __ldcg(src + (size_t)bb * H) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    ax += v[u].x;
                    ay = has_max ? fmax(ay, v[u].y) : ay + v[u].y;
                }
            }
            fold[g][2 * h] = ax;
            fold[g][2 * h + 1] = ay;
        }
    }
    __syncthreads();
    if (threadIdx.x < NSUM) {
        const int q = threadIdx.x;
        double a = fold[0][q];
        for (int g = 1; g < G; ++g) a = (q == NSUM - 1) ? fmax(a, fold[g][q]) : a + fold[g][q];
        folded[q] = a;
    }
    __syncthreads();
    Sums acc;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NSUM; ++q) acc.v[q] = folded[q];
    }
    const unsigned long long t_last = gtime();  // every block has finished its atoms
    if (mode & FIN_P2P) {
        // All-gather of the rank sums through peer memory, fused into this kernel: every rank stores its 12 sums into every
        // rank's mailbox (NVLink stores), raises its sequence flag there, waits for the flags of all ranks in its own
        // mailbox and folds the ranks in rank order — identical lambda / myu / rebuild decision everywhere.
        __shared__ double my_sums[NSUM];
        __shared__ int timed_out;
        const Peers &peers = *peers_p;
        const unsigned long long seq = sc->epoch + 1;
        const int buf = (int)(seq & 1ull);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int q = 0; q < NSUM; ++q) my_sums[q] = acc.v[q];
            timed_out = 0;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < peers.nranks * NSUM; t += BLOCK) {
            const int r = t / NSUM, q = t - r * NSUM;
            *reinterpret_cast<volatile double *>(&peers.mail[r]->sums[buf][peers.rank][q]) = my_sums[q];
        }
        __syncthreads();  // the stores above happen-before the release stores below (cumulative over the barrier)
        const unsigned long long t_wait = gtime();
        if ((int)threadIdx.x < peers.nranks) {
            st_release_sys(&peers.mail[threadIdx.x]->sums_seq[peers.rank], seq);
            if (!wait_seq(&peers.mail[peers.rank]->sums_seq[threadIdx.x], seq)) timed_out = 1;
        }
        __syncthreads();
        if (threadIdx.x == 0) sc->wait_sums_ns += gtime() - t_wait;
        if (threadIdx.x == 0) {
            Sums t;
#pragma unroll
            for (int q = 0; q < NSUM; ++q) t.v[q] = 0.0;
            const Mail *own = peers.mail[peers.rank];
            for (int r = 0; r < peers.nranks; ++r) {
#pragma unroll
                for (int q = 0; q < NSUM - 1; ++q) t.v[q] += __ldcg(&own->sums[buf][r][q]);
                t.v[NSUM - 1] = fmax(t.v[NSUM - 1], __ldcg(&own->sums[buf][r][NSUM - 1]));
            }
            finalize(sc, &sc_in, &pr_in, t, mode);
            if (timed_out) sc->error = 3;  // MD_ERR_NCCL: a peer never delivered
            sc->force_atoms_ns += t_last - sc->t_start;
            sc->force_tail_ns += gtime() - t_last;
            sc->t_start = ~0ull;
            sc->epoch = seq;
            sc->ticket = 0;
            if (cond_handle) {
                unsigned int go = (sc->steps_left > 0 && !sc->need_rebuild && !sc->error) ? 1u : 0u;
                cudaGraphSetConditional((cudaGraphConditionalHandle)cond_handle, go);
            }
        }
        return;
    }
    if (threadIdx.x == 0) {
        PROBE(3);
        if (mode & FIN_DIST) {
#pragma unroll
            for (int q = 0; q < NSUM; ++q) sc->rank_sums[q] = acc.v[q];
            sc->ticket = 0;
            return;
        }
        finalize(sc, &sc_in, &pr_in, acc, mode);
        sc->ticket = 0;
        // everything above is visible to whoever acquires the new sequence number (early-start k_kick_drift)
        st_release_gpu(&sc->fin_seq, sc_in.fin_seq + 1);
        PROBE(4);
        if (cond_handle) {
            unsigned int go = (sc->steps_left > 0 && !sc->need_rebuild && !sc->error) ? 1u : 0u;
            cudaGraphSetConditional((cudaGraphConditionalHandle)cond_handle, go);
        }
    }
}

// Adds one atom's terms. (wx,wy,wz) = v + F*c is the velocity the next kick_drift moves this atom with (before lambda).
__device__ __forceinline__ void accumulate_sums(Sums &s, double m, double vx, double vy, double vz, double wx,
                                                double wy, double wz, double w, double u, const double *shift)
{
    s.v[0] += m * vx; s.v[1] += m * vy; s.v[2] += m * vz;
    double ax = vx - shift[0], ay = vy - shift[1], az = vz - shift[2];
    s.v[S_TH] += m * (ax * ax + ay * ay + az * az);
    s.v[S_KE] += m * (vx * vx + vy * vy + vz * vz);
    s.v[S_W] += w;
    s.v[S_U] += u;
    s.v[S_MU] += m * wx; s.v[S_MU + 1] += m * wy; s.v[S_MU + 2] += m * wz;
    double bx = wx - shift[0], by = wy - shift[1], bz = wz - shift[2];
    s.v[S_THU] += m * (bx * bx + by * by + bz * bz);
    s.v[S_MAX] = fmax(s.v[S_MAX], wx * wx + wy * wy + wz * wz);
}

// Standalone K5 over the stored state (after upload, or when only the macro parameters are wanted).
constexpr int RED_BLOCK = 256;
__global__ void __launch_bounds__(RED_BLOCK) k_reduce_state(int n, Arrays a, double *__restrict__ partials,
                                                            Scalars *sc, const Params *__restrict__ pr, int mode)
{
    Sums s;
#pragma unroll
    for (int q = 0; q < NSUM; ++q) s.v[q] = 0.0;
    const double shift[3] = {sc->shift[0], sc->shift[1], sc->shift[2]};
    const double c = pr->half_dt_m, m = pr->mass;
    for (int i = blockIdx.x * RED_BLOCK + threadIdx.x; i < n; i += gridDim.x * RED_BLOCK) {
        double vx = a.vx[i], vy = a.vy[i], vz = a.vz[i];
        double wx = __dadd_rn(vx, __dmul_rn(a.fx[i], c)), wy = __dadd_rn(vy, __dmul_rn(a.fy[i], c)),
               wz = __dadd_rn(vz, __dmul_rn(a.fz[i], c));
        accumulate_sums(s, m, vx, vy, vz, wx, wy, wz, a.w[i], a.u[i], shift);
    }
    block_reduce<RED_BLOCK>(s);
    grid_reduce_finalize<RED_BLOCK>(s, partials, sc, pr, mode, 0ull, nullptr);
}

// ----------------------------------------------------------------------------------------------------
// K3: pair forces from the Verlet list (each ordered pair evaluated from both sides, like the reference — no
// Newton-3 sharing, no atomics, deterministic), fused with both half-kicks that surround it and the K5 sums.
//   EXACT: potential.rs:181-211 operation by operation, no FMA, partners in ascending upload index.
//   FAST : r²-based Lennard-Jones (one division, no sqrt), FMA allowed.
// Persistent grid (a fixed number of blocks, grid-stride over atoms): few per-block partials for the final
// fixed-order reduction, and the atom→thread assignment (hence every sum) is fixed for a given grid.
//
// Velocity planes: on entry of a step they hold u = v + F_old*c (first half-kick done, thermostat scale not yet).
//   v'  = lambda * u                       thermostat.rs:54-58 (the same product k_kick_drift drifted with)
//   v'' = v' + F*c                         integrator.rs:47-53 — end-of-step velocity, enters the K5 sums
//   u'  = v'' + F*c                        integrator.rs:28-34 of the NEXT step (same F, same c)
// Steady state stores only u' (72 B/atom in, 24 B/atom out); the last step of a batch stores v'', F, U, W so the
// resident State is complete whenever the host can observe it.
#ifndef MD_FORCE_MINB
#define MD_FORCE_MINB 4
#endif
#ifndef MD_FORCE_MINB_DILUTE
#define MD_FORCE_MINB_DILUTE 4
#endif
#ifndef MD_DILUTE_ROWS
#define MD_DILUTE_ROWS 1
#endif
constexpr int FORCE_BLOCK = 128;

// Launch constants of the force kernel: passed BY VALUE so they live in the constant bank and feed FP64 instructions
// as c[bank][offset] operands instead of occupying ~20 registers per thread.
struct ForceConsts {
    double sigma, sigma2, eps4, eps24, r_cut, rc2, u_cut;
    double c6, c12, d6, d12;  // 24 eps sigma^6, 48 eps sigma^12, 4 eps sigma^6, 4 eps sigma^12 (dense FAST pair term)
    double hc;    // dt / (2 m)
    double mass;
};

struct LjConst {
    double Lx, Ly, Lz, hx, hy, hz;
    int hxi, hyi, hzi;  // high words of hx, hy, hz: integer-pipe pre-test of the minimum-image condition
};

struct PairAcc {
    double fx, fy, fz, u, w;
};

// |r| >= h can only hold if the high word of |r| is >= the high word of h: the common (no wrap) case costs one
// integer compare instead of two FP64 compares; the exact single-shift rule runs only when the pre-test fires.
__device__ __forceinline__ double min_image_fast(double r, double L, double h, int hhi)
{
    if ((__double2hiint(r) & 0x7fffffff) >= hhi) r = min_image(r, L, h);
    return r;
}

// 1/x without the IEEE division's slow-path branch: MUFU.RCP64H seed (relative error <= 2^-23) + one Newton step →
// <= 2^-46 (1.4e-14), three orders below the 1e-10 parity bar of the FAST mode.
__device__ __forceinline__ double rcp_nr(double x)
{
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double e = fma(-x, y, 1.0);
    return fma(y, e, y);
}

// A load the compiler can neither hoist nor keep live across a loop: per-block constants that are only needed between two
// long neighbour loops are re-read (L1/L2 hits) instead of occupying registers inside them.
__device__ __forceinline__ double ld_pinned(const double *p)
{
    double v;
    asm volatile("ld.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// branch-free single-shift minimum image (same rule as min_image): two compares, a select, one add
__device__ __forceinline__ double min_image_sel(double r, double L, double h)
{
    const double s = r > h ? -L : (r < -h ? L : 0.0);
    return r + s;
}

// FAST pair term for dense systems, branch-free: masked pairs (k beyond this atom's list, or outside the cutoff)
// contribute exact zeros.  WRAP = false is used by warps whose atoms all sit further than r_list + skin from every box
// face: none of their partners can be a periodic image, so the minimum-image step is skipped altogether.
template <bool WRAP>
__device__ __forceinline__ void pair_fast(PairAcc &a, bool active, double xj, double yj, double zj, double xi,
                                          double yi, double zi, const LjConst &c, const ForceConsts &fc, bool need_u,
                                          bool need_w)
{
    double rx = xj - xi, ry = yj - yi, rz = zj - zi;
    if (WRAP) {
        rx = min_image_sel(rx, c.Lx, c.hx);
        ry = min_image_sel(ry, c.Ly, c.hy);
        rz = min_image_sel(rz, c.Lz, c.hz);
    }
    double r2 = rx * rx + ry * ry + rz * rz;
    bool in = active && (r2 <= fc.rc2);
    // r2 > 0 for every lane: masked lanes gather an atom that is not one of the thread's own (see safe_dummy), so the
    // reciprocal needs no guard — whatever it yields for a masked or out-of-range pair is discarded by the select below
    // with y = 1/r^2:  F/r = 24 eps (s^6 - 2 s^12) / r^2 = y^4 (c6 - c12 y^3),  U = y^3 (d12 y^3 - d6) - u_cut
    const double y = rcp_nr(r2);
    const double y2 = y * y;
    const double y3 = y2 * y;
    double fr = (y2 * y2) * fma(-fc.c12, y3, fc.c6);  // F / r
    fr = in ? fr : 0.0;
    a.fx += fr * rx; a.fy += fr * ry; a.fz += fr * rz;
    // per-atom potential / virial: uniform flags — steady-state steps of a batch only need what feeds the controls
    if (need_u) {
        const double pu = fma(y3, fma(fc.d12, y3, -fc.d6), -fc.u_cut);
        a.u += in ? pu : 0.0;
    }
    if (need_w) a.w += fr * r2;
}

// FAST pair term for dilute systems: most listed partners are outside the cutoff (the skin is wide), so the
// Lennard-Jones body sits behind a real branch and the FP64 pipe only sees the cheap distance test.
__device__ __forceinline__ void pair_fast_branchy(PairAcc &a, bool active, double xj, double yj, double zj,
                                                  double xi, double yi, double zi, const LjConst &c,
                                                  const ForceConsts &fc)
{
    double rx = min_image_fast(xj - xi, c.Lx, c.hx, c.hxi);
    double ry = min_image_fast(yj - yi, c.Ly, c.hy, c.hyi);
    double rz = min_image_fast(zj - zi, c.Lz, c.hz, c.hzi);
    double r2 = rx * rx + ry * ry + rz * rz;
    if (active && r2 <= fc.rc2) {
        double inv = 1.0 / r2;
        double s2 = fc.sigma2 * inv;
        double s6 = s2 * s2 * s2;
        double s12 = s6 * s6;
        double fr = fc.eps24 * inv * (s6 - 2.0 * s12);
        a.u += fc.eps4 * (s12 - s6) - fc.u_cut;
        a.fx += fr * rx; a.fy += fr * ry; a.fz += fr * rz;
        a.w += fr * r2;
    }
}

// EXACT pair term: potential.rs:181-211 operation by operation, no contraction, real branch on the cutoff.
__device__ __forceinline__ void pair_exact(PairAcc &a, double xj, double yj, double zj, double xi, double yi,
                                           double zi, const LjConst &c, const ForceConsts &fc)
{
    double rx = min_image(__dsub_rn(xj, xi), c.Lx, c.hx);
    double ry = min_image(__dsub_rn(yj, yi), c.Ly, c.hy);
    double rz = min_image(__dsub_rn(zj, zi), c.Lz, c.hz);
    double r = norm_exact(rx, ry, rz);
    if (r > fc.r_cut) return;                            // potential.rs:202 (inclusive cutoff)
    double sr = __ddiv_rn(fc.sigma, r);                   // potential.rs:63
    double x2 = __dmul_rn(sr, sr), x4 = __dmul_rn(x2, x2);
    double s6 = __dmul_rn(x2, x4);                        // powi(6) = x² · x⁴
    double s12 = __dmul_rn(s6, s6);
    double pu = __dsub_rn(__dmul_rn(fc.eps4, __dsub_rn(s12, s6)), fc.u_cut);
    double pf = __dmul_rn(__ddiv_rn(fc.eps24, r), __dsub_rn(s6, __dmul_rn(2.0, s12)));
    double vx = __dmul_rn(__ddiv_rn(rx, r), pf);          // r / r_abs * force   potential.rs:207
    double vy = __dmul_rn(__ddiv_rn(ry, r), pf);
    double vz = __dmul_rn(__ddiv_rn(rz, r), pf);
    double t = __dadd_rn(__dadd_rn(__dmul_rn(vx, rx), __dmul_rn(vy, ry)), __dmul_rn(vz, rz));
    a.fx = __dadd_rn(a.fx, vx); a.fy = __dadd_rn(a.fy, vy); a.fz = __dadd_rn(a.fz, vz);
    a.u = __dadd_rn(a.u, pu);
    a.w = __dadd_rn(a.w, t);
}

// Both half-kicks around the force (see header comment above), the K5 terms, and the stores of one atom.
// per-thread running sums kept in shared memory (column per thread → conflict-free), not in 16 registers
struct SumsSmem {
    double v[NSUM][FORCE_BLOCK];
};

// nh (uniform): also accumulate the COM/thermal sums of u', which only Nose-Hoover's second psi update reads.
__device__ __forceinline__ void finish_atom(SumsSmem &ss, const PairAcc &f, double &vx, double &vy, double &vz,
                                            bool do_step, double lambda, double c, double mass, const double *shift,
                                            double &wx, double &wy, double &wz, bool nh)
{
    if (do_step) {
        vx = __dadd_rn(__dmul_rn(vx, lambda), __dmul_rn(f.fx, c));  // v'' = lambda*u + F*c
        vy = __dadd_rn(__dmul_rn(vy, lambda), __dmul_rn(f.fy, c));
        vz = __dadd_rn(__dmul_rn(vz, lambda), __dmul_rn(f.fz, c));
    }
    wx = __dadd_rn(vx, __dmul_rn(f.fx, c));                         // u' = v'' + F*c
    wy = __dadd_rn(vy, __dmul_rn(f.fy, c));
    wz = __dadd_rn(vz, __dmul_rn(f.fz, c));
    const int l = threadIdx.x;
    ss.v[0][l] += mass * vx; ss.v[1][l] += mass * vy; ss.v[2][l] += mass * vz;
    const double ax = vx - shift[0], ay = vy - shift[1], az = vz - shift[2];
    ss.v[3][l] += mass * (ax * ax + ay * ay + az * az);
    ss.v[4][l] += mass * (vx * vx + vy * vy + vz * vz);
    ss.v[S_W][l] += f.w;
    ss.v[S_U][l] += f.u;
    if (nh) {
        ss.v[S_MU][l] += mass * wx; ss.v[S_MU + 1][l] += mass * wy; ss.v[S_MU + 2][l] += mass * wz;
        const double bx = wx - shift[0], by = wy - shift[1], bz = wz - shift[2];
        ss.v[S_THU][l] += mass * (bx * bx + by * by + bz * bz);
    }
    ss.v[S_MAX][l] = fmax(ss.v[S_MAX][l], wx * wx + wy * wy + wz * wz);
}

// Neighbour loop of one atom pair (FAST modes).  The next rows of partner indices are prefetched while the current
// ones are in flight; MASKED = branch-free pair term + packed gathers (dense), else branchy pair term + plane gathers.
template <int ROWS, bool MASKED, bool WRAP>
__device__ __forceinline__ void neighbour_loop(PairAcc &f0, PairAcc &f1, const Arrays &a, const int2 *__restrict__ row,
                                               size_t stride, int last_row, int2 C, int i0, double2 X, double2 Y,
                                               double2 Z, const LjConst &c, const ForceConsts &fc, int2 Ja)
{
    // Ja = row[0]: it exists for every atom (cap >= 8) and the caller fetched it together with the atom's own data
    const double *__restrict__ px = a.x, *__restrict__ py = a.y, *__restrict__ pz = a.z;
    const int kmax = max(C.x, C.y);
    int k = 0;
    if (ROWS == 2) {
        int2 Jb = row[min(1, last_row) * stride];
        for (; k + 1 < kmax; k += 2) {
            const int2 Na = row[min(k + 2, last_row) * stride], Nb = row[min(k + 3, last_row) * stride];
            const bool a0 = k < C.x, a1 = k < C.y, b0 = k + 1 < C.x, b1 = k + 1 < C.y;
            const int ja0 = a0 ? Ja.x : i0, ja1 = a1 ? Ja.y : i0, jb0 = b0 ? Jb.x : i0, jb1 = b1 ? Jb.y : i0;
            double xa0, ya0, za0, xa1, ya1, za1, xb0, yb0, zb0, xb1, yb1, zb1;
            if (MASKED) {  // dense: one 32 B sector per partner from the packed copy
                const double4 qa0 = a.q4[ja0], qa1 = a.q4[ja1], qb0 = a.q4[jb0], qb1 = a.q4[jb1];
                xa0 = qa0.x; ya0 = qa0.y; za0 = qa0.z; xa1 = qa1.x; ya1 = qa1.y; za1 = qa1.z;
                xb0 = qb0.x; yb0 = qb0.y; zb0 = qb0.z; xb1 = qb1.x; yb1 = qb1.y; zb1 = qb1.z;
                pair_fast<WRAP>(f0, a0, xa0, ya0, za0, X.x, Y.x, Z.x, c, fc, true, true);
                pair_fast<WRAP>(f1, a1, xa1, ya1, za1, X.y, Y.y, Z.y, c, fc, true, true);
                pair_fast<WRAP>(f0, b0, xb0, yb0, zb0, X.x, Y.x, Z.x, c, fc, true, true);
                pair_fast<WRAP>(f1, b1, xb1, yb1, zb1, X.y, Y.y, Z.y, c, fc, true, true);
            } else {
                xa0 = px[ja0]; ya0 = py[ja0]; za0 = pz[ja0]; xa1 = px[ja1]; ya1 = py[ja1]; za1 = pz[ja1];
                xb0 = px[jb0]; yb0 = py[jb0]; zb0 = pz[jb0]; xb1 = px[jb1]; yb1 = py[jb1]; zb1 = pz[jb1];
                pair_fast_branchy(f0, a0, xa0, ya0, za0, X.x, Y.x, Z.x, c, fc);
                pair_fast_branchy(f1, a1, xa1, ya1, za1, X.y, Y.y, Z.y, c, fc);
                pair_fast_branchy(f0, b0, xb0, yb0, zb0, X.x, Y.x, Z.x, c, fc);
                pair_fast_branchy(f1, b1, xb1, yb1, zb1, X.y, Y.y, Z.y, c, fc);
            }
            Ja = Na; Jb = Nb;
        }
    }
    for (; k < kmax; ++k) {
        const int2 Na = row[min(k + 1, last_row) * stride];
        const bool a0 = k < C.x, a1 = k < C.y;
        const int ja0 = a0 ? Ja.x : i0, ja1 = a1 ? Ja.y : i0;
        if (MASKED) {
            const double4 qa0 = a.q4[ja0], qa1 = a.q4[ja1];
            pair_fast<WRAP>(f0, a0, qa0.x, qa0.y, qa0.z, X.x, Y.x, Z.x, c, fc, true, true);
            pair_fast<WRAP>(f1, a1, qa1.x, qa1.y, qa1.z, X.y, Y.y, Z.y, c, fc, true, true);
        } else {
            const double xa0 = px[ja0], ya0 = py[ja0], za0 = pz[ja0];
            const double xa1 = px[ja1], ya1 = py[ja1], za1 = pz[ja1];
            pair_fast_branchy(f0, a0, xa0, ya0, za0, X.x, Y.x, Z.x, c, fc);
            pair_fast_branchy(f1, a1, xa1, ya1, za1, X.y, Y.y, Z.y, c, fc);
        }
        Ja = Na;
    }
}

// ---- dense pair term ---------------------------------------------------------------------------------------------------
// What the SASS of the first dense loop showed (cuobjdump, 4 pair terms per trip: 227 instructions, 84 of them FP64): the
// uniform need_u / need_w flags had been if-converted — potential and virial were computed for every pair and dropped by a
// select (16 FP64 + 12 FSEL per trip).  The flags are therefore a template parameter (UW: 0 = forces only, 1 = + virial,
// 2 = + potential); one of six loop instances runs per launch.  Same arithmetic in the same order: bit-identical results.

// single-shift minimum image, same rule as min_image (r > h → r - L, r < -h → r + L) written as |r| > h → r - copysign(L, r):
// one FP64 compare instead of two, the sign work on the integer pipe
__device__ __forceinline__ double min_image_abs(double r, double L, double h)
{
    const int hi = __double2hiint(r);
    const double ar = __hiloint2double(hi & 0x7fffffff, __double2loint(r));
    const double s = __hiloint2double(__double2hiint(L) | (~hi & 0x80000000), __double2loint(L));  // -copysign(L, r), L > 0
    return r + (ar > h ? s : 0.0);
}

template <bool WRAP, int UW>
__device__ __forceinline__ void pair_dense(PairAcc &a, bool active, double xj, double yj, double zj, double xi, double yi,
                                           double zi, const LjConst &c, const ForceConsts &fc)
{
    double rx = xj - xi, ry = yj - yi, rz = zj - zi;
    if (WRAP) {
        rx = min_image_abs(rx, c.Lx, c.hx);
        ry = min_image_abs(ry, c.Ly, c.hy);
        rz = min_image_abs(rz, c.Lz, c.hz);
    }
    const double r2 = rx * rx + ry * ry + rz * rz;
    const bool in = active && (r2 <= fc.rc2);
    // see pair_fast: r2 > 0 on every lane, masked and out-of-range pairs are zeroed by the select
    const double y = rcp_nr(r2);
    const double y2 = y * y;
    const double y3 = y2 * y;
    double fr = (y2 * y2) * fma(-fc.c12, y3, fc.c6);  // F / r
    fr = in ? fr : 0.0;
    a.fx += fr * rx; a.fy += fr * ry; a.fz += fr * rz;
    if (UW >= 2) {
        const double pu = fma(y3, fma(fc.d12, y3, -fc.d6), -fc.u_cut);
        a.u += in ? pu : 0.0;
    }
    if (UW >= 1) a.w += fr * r2;
}

// Dense systems (hundreds of listed partners per atom): the neighbour table is far larger than L2 and streams from HBM, so a
// one-trip-ahead index prefetch leaves the warp waiting on DRAM every trip.  Each thread therefore keeps a ring of the next
// RING_D trips' index rows (two rows per trip) in shared memory, filled by cp.async — no registers, no barrier (a thread only
// reads what it copied), ~RING_D trips of DRAM latency hidden.
// Address masked lanes gather from: warp-uniform (one L1 pass), a real atom, and never one of the warp's own 64 atoms — so
// its distance to the lane's atoms is positive and the pair term stays finite before it is masked out.  (n >= 128 on this path.)
__device__ __forceinline__ int safe_dummy(int i0, int n)
{
    const int w0 = i0 & ~63;
    return w0 + 64 < n ? w0 + 64 : w0 - 64;
}

constexpr int RING_D = 8;
struct IndexRing {
    int2 r[RING_D][2][FORCE_BLOCK];
};

template <bool WRAP, int UW>
__device__ __forceinline__ void neighbour_loop_dense(PairAcc &f0, PairAcc &f1, const double4 *__restrict__ q4,
                                                     const int2 *__restrict__ row, size_t stride, int2 C, int i0,
                                                     double2 X, double2 Y, double2 Z, const LjConst &c,
                                                     const ForceConsts &fc, IndexRing &ring, int n)
{
    const int l = threadIdx.x;
    const int kmax = max(C.x, C.y);
    const int ntrips = (kmax + 1) >> 1;
    // rows beyond this pair's lists are never read (the copies are skipped, the lanes masked)
#pragma unroll
    for (int d = 0; d < RING_D; ++d) {
        if (d < ntrips) {
            cp_async8(&ring.r[d][0][l], row + (size_t)(2 * d) * stride);
            if (2 * d + 1 < kmax) cp_async8(&ring.r[d][1][l], row + (size_t)(2 * d + 1) * stride);
        }
        cp_async_commit();
    }
    const int2 *refill = row + (size_t)(2 * RING_D) * stride;
    // software pipeline: the gathers of trip t+1 are in flight while the pair terms of trip t are computed.
#define MD_FETCH_ROWS(T, JA, JB)                                                                           \
    do {                                                                                                   \
        const int slot_ = (T) % RING_D;                                                                    \
        asm volatile("cp.async.wait_group %0;" ::"n"(RING_D - 1) : "memory");                              \
        JA = ring.r[slot_][0][l];                                                                          \
        JB = ring.r[slot_][1][l];                                                                          \
        const int tn_ = (T) + RING_D; /* refill the slot with the rows of trip T + RING_D */               \
        if (tn_ < ntrips) {                                                                                \
            cp_async8(&ring.r[slot_][0][l], refill);                                                       \
            if (2 * tn_ + 1 < kmax) cp_async8(&ring.r[slot_][1][l], refill + stride);                      \
        }                                                                                                  \
        cp_async_commit();                                                                                 \
        refill += 2 * stride; /* walks the table two rows per trip: no 64-bit multiply per refill */       \
    } while (0)
    // one 256-bit load per partner (LDG.E.256, new with sm_100): a divergent gather costs the L1 one pass per lane and
    // instruction, and this loop is co-limited by exactly that — half the passes of an (x, y) + z pair of loads
#define MD_GATHER(J, XY, ZZ)                                                                              \
    do {                                                                                                  \
        double w_;                                                                                        \
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"                                                    \
            : "=d"(XY.x), "=d"(XY.y), "=d"(ZZ), "=d"(w_)                                                  \
            : "l"(q4 + (J)));                                                                             \
    } while (0)
    // masked lanes (list shorter than the warp's longest) all gather the same address: one L1 pass instead of 32
    i0 = safe_dummy(i0, n);
    double2 pa0, pa1, pb0, pb1;
    double za0, za1, zb0, zb1;
    pa0 = pa1 = pb0 = pb1 = make_double2(0.0, 0.0);
    za0 = za1 = zb0 = zb1 = 0.0;
    if (ntrips > 0) {
        int2 Ja, Jb;
        MD_FETCH_ROWS(0, Ja, Jb);
        MD_GATHER(0 < C.x ? Ja.x : i0, pa0, za0); MD_GATHER(0 < C.y ? Ja.y : i0, pa1, za1);
        MD_GATHER(1 < C.x ? Jb.x : i0, pb0, zb0); MD_GATHER(1 < C.y ? Jb.y : i0, pb1, zb1);
    }
    for (int t = 0; t < ntrips; ++t) {
        const int k = 2 * t;
        double2 na0 = pa0, na1 = pa1, nb0 = pb0, nb1 = pb1;
        double ya0 = za0, ya1 = za1, yb0 = zb0, yb1 = zb1;
        if (t + 1 < ntrips) {
            int2 Ja, Jb;
            MD_FETCH_ROWS(t + 1, Ja, Jb);
            MD_GATHER(k + 2 < C.x ? Ja.x : i0, na0, ya0); MD_GATHER(k + 2 < C.y ? Ja.y : i0, na1, ya1);
            MD_GATHER(k + 3 < C.x ? Jb.x : i0, nb0, yb0); MD_GATHER(k + 3 < C.y ? Jb.y : i0, nb1, yb1);
        }
        const bool a0 = k < C.x, a1 = k < C.y, b0 = k + 1 < C.x, b1 = k + 1 < C.y;
        pair_dense<WRAP, UW>(f0, a0, pa0.x, pa0.y, za0, X.x, Y.x, Z.x, c, fc);
        pair_dense<WRAP, UW>(f1, a1, pa1.x, pa1.y, za1, X.y, Y.y, Z.y, c, fc);
        pair_dense<WRAP, UW>(f0, b0, pb0.x, pb0.y, zb0, X.x, Y.x, Z.x, c, fc);
        pair_dense<WRAP, UW>(f1, b1, pb1.x, pb1.y, zb1, X.y, Y.y, Z.y, c, fc);
        pa0 = na0; pa1 = na1; pb0 = nb0; pb1 = nb1;
        za0 = ya0; za1 = ya1; zb0 = yb0; zb1 = yb1;
    }
#undef MD_FETCH_ROWS
#undef MD_GATHER
    cp_async_wait_all();
}

// Dense systems with UNION lists (k_build_union): one entry = one gather, evaluated against both atoms of the thread under
// the entry's membership bits.  Same ring / pipeline structure as neighbour_loop_dense, half the gathers per pair term.
template <bool WRAP, int UW>
__device__ __forceinline__ void neighbour_loop_union(PairAcc &f0, PairAcc &f1, const double4 *__restrict__ q4,
                                                     const int *__restrict__ row, size_t stride, int cnt, int i0,
                                                     double2 X, double2 Y, double2 Z, const LjConst &c,
                                                     const ForceConsts &fc, IndexRing &ring, int n)
{
    const int l = threadIdx.x;
    const int ntrips = (cnt + 1) >> 1;
    int *slots = reinterpret_cast<int *>(&ring.r[0][0][0]);  // [RING_D][2][FORCE_BLOCK] ints
#define MD_SLOT(D, H) slots[((D) * 2 + (H)) * FORCE_BLOCK + l]
#define MD_CP4(DST, SRC) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(DST)), "l"(SRC) : "memory")
#pragma unroll
    for (int d = 0; d < RING_D; ++d) {
        if (d < ntrips) {
            MD_CP4(&MD_SLOT(d, 0), row + (size_t)(2 * d) * stride);
            if (2 * d + 1 < cnt) MD_CP4(&MD_SLOT(d, 1), row + (size_t)(2 * d + 1) * stride);
        }
        cp_async_commit();
    }
    i0 = safe_dummy(i0, n);  // masked entries gather one common address
#define MD_FETCH_ENTRIES(T, EA, EB)                                                                   \
    do {                                                                                              \
        const int slot_ = (T) % RING_D;                                                               \
        asm volatile("cp.async.wait_group %0;" ::"n"(RING_D - 1) : "memory");                         \
        EA = MD_SLOT(slot_, 0);                                                                       \
        EB = 2 * (T) + 1 < cnt ? MD_SLOT(slot_, 1) : i0;                                              \
        const int tn_ = (T) + RING_D;                                                                 \
        if (tn_ < ntrips) {                                                                           \
            MD_CP4(&MD_SLOT(slot_, 0), row + (size_t)(2 * tn_) * stride);                             \
            if (2 * tn_ + 1 < cnt) MD_CP4(&MD_SLOT(slot_, 1), row + (size_t)(2 * tn_ + 1) * stride);  \
        }                                                                                             \
        cp_async_commit();                                                                            \
    } while (0)
#define MD_GATHER(J, XY, ZZ)                                                                              \
    do {                                                                                                  \
        double w_;                                                                                        \
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"                                                    \
            : "=d"(XY.x), "=d"(XY.y), "=d"(ZZ), "=d"(w_)                                                  \
            : "l"(q4 + (J)));                                                                             \
    } while (0)
    double2 pa = make_double2(0.0, 0.0), pb = pa;
    double za = 0.0, zb = 0.0;
    int ea = i0, eb = i0;
    if (ntrips > 0) {
        MD_FETCH_ENTRIES(0, ea, eb);
        MD_GATHER(ea & UNION_IDX, pa, za);
        MD_GATHER(eb & UNION_IDX, pb, zb);
    }
    for (int t = 0; t < ntrips; ++t) {
        double2 na = pa, nb = pb;
        double ya = za, yb = zb;
        int fa = i0, fb = i0;
        if (t + 1 < ntrips) {
            MD_FETCH_ENTRIES(t + 1, fa, fb);
            MD_GATHER(fa & UNION_IDX, na, ya);
            MD_GATHER(fb & UNION_IDX, nb, yb);
        }
        pair_dense<WRAP, UW>(f0, (ea & UNION_A) != 0, pa.x, pa.y, za, X.x, Y.x, Z.x, c, fc);
        pair_dense<WRAP, UW>(f1, ea < 0, pa.x, pa.y, za, X.y, Y.y, Z.y, c, fc);
        pair_dense<WRAP, UW>(f0, (eb & UNION_A) != 0, pb.x, pb.y, zb, X.x, Y.x, Z.x, c, fc);
        pair_dense<WRAP, UW>(f1, eb < 0, pb.x, pb.y, zb, X.y, Y.y, Z.y, c, fc);
        pa = na; pb = nb; za = ya; zb = yb; ea = fa; eb = fb;
    }
#undef MD_FETCH_ENTRIES
#undef MD_GATHER
#undef MD_SLOT
#undef MD_CP4
    cp_async_wait_all();
}

// ----------------------------------------------------------------------------------------------------
// K3 for dilute systems, FAST mode: k_force_sparse (opt-in experiment, MOLDYN_B200_SPARSE=1 — slower than k_force on B200).
// In a 300 K argon gas ~78 % of the atoms have NO listed partner (even with skin = r_cut), but with two atoms per thread
// and 32 threads per warp every warp of k_force still walks the whole gather path with most lanes idle.  Here the work
// is split by atom class, inside one launch and with the same persistent grid:
//   phase S  every atom WITHOUT partners: F = 0, so neither its position nor the list is read — 28 B in (u, count),
//            24 B out per atom, two independent pairs of atoms in flight per thread;
//   phase A  the atoms WITH partners, through the compacted index list built at the last rebuild (k_flag_active + scan):
//            one atom per thread, every lane has real gather work.
// The arithmetic of an atom is the same finish_atom() as everywhere else (with F = 0 for phase S), the atom → thread map is
// fixed by the grid, so results stay run-to-run reproducible.
__global__ void k_flag_active(int n, const int *__restrict__ nbr_cnt, int *__restrict__ flag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = nbr_cnt[i] > 0 ? 1 : 0;
}

// shared by the force kernels: guarded early-out, phase clock, wait for the neighbours' ghosts (peer-memory path)
__device__ __forceinline__ bool force_prologue(int do_step, Scalars *sc, const Peers *peers)
{
    if (do_step & 32) { pdl_wait(); pdl_launch_dependents(); }
    if ((do_step & 4) && halted(sc)) return false;  // uniform over the grid: nobody takes a ticket
    if ((do_step & 8) && threadIdx.x == 0) atomicMin(&sc->t_start, gtime());
    if (do_step & 16) {
        __shared__ int halo_late;
        if (threadIdx.x == 0) {
            const unsigned long long seq = sc->epoch + 1;
            // Our own face atoms were stored into the neighbours' planes by the preceding k_kick_drift; a kernel boundary
            // orders those stores system-wide, so the flags can go up right away — no fence inside the drift kernel.
            if (blockIdx.x == 0) {
                st_release_sys(&peers->mail[peers->left]->halo_seq[1], seq);   // we are the left neighbour's right side
                st_release_sys(&peers->mail[peers->right]->halo_seq[0], seq);
            }
            const Mail *own = peers->mail[peers->rank];
            const unsigned long long t0 = gtime();
            halo_late = !(wait_seq(&own->halo_seq[0], seq) && wait_seq(&own->halo_seq[1], seq));
            if (blockIdx.x == 0) sc->wait_halo_ns += gtime() - t0;
        }
        __syncthreads();
        if (halo_late && threadIdx.x == 0) atomicExch(&sc->error, 3);
    }
    return true;
}

__global__ void __launch_bounds__(FORCE_BLOCK, 5)
    k_force_sparse(int n, Arrays a, const int *__restrict__ nbr, const int *__restrict__ nbr_cnt, int npad,
                   const int *__restrict__ active_idx, const int *__restrict__ n_active_p, double *__restrict__ partials,
                   Scalars *sc, const Params *__restrict__ pr, int do_step, unsigned long long cond_handle,
                   const ForceConsts fc, const Peers *peers)
{
    if (!force_prologue(do_step, sc, peers)) return;
    __shared__ SumsSmem ss;
#pragma unroll
    for (int q = 0; q < NSUM; ++q) ss.v[q][threadIdx.x] = 0.0;
    const bool step = (do_step & 1) != 0;
    const bool store_state = !step || sc->steps_left <= 1;
    const bool nh = pr->th_kind == 2 || !step;
    const double lambda = sc->lambda;
    const double shift[3] = {sc->shift[0], sc->shift[1], sc->shift[2]};
    const PairAcc zero = {0.0, 0.0, 0.0, 0.0, 0.0};
    const int tstride = gridDim.x * FORCE_BLOCK;

    // ---- phase S: atoms without partners ---------------------------------------------------------------------------
    const int npairs = (n + 1) >> 1;
    for (int t0 = blockIdx.x * FORCE_BLOCK + threadIdx.x; t0 < npairs; t0 += 2 * tstride) {
        // two pairs of atoms per trip: all loads first
        const int t1 = t0 + tstride;
        const bool two = t1 < npairs;
        int2 C[2];
        double2 VX[2], VY[2], VZ[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int t = u ? t1 : t0;
            if (u == 0 || two) {
                C[u] = reinterpret_cast<const int2 *>(nbr_cnt)[t];
                VX[u] = reinterpret_cast<const double2 *>(a.vx)[t]; VY[u] = reinterpret_cast<const double2 *>(a.vy)[t];
                VZ[u] = reinterpret_cast<const double2 *>(a.vz)[t];
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int t = u ? t1 : t0;
            if (u == 1 && !two) break;
            const int i0 = 2 * t;
            const bool has1 = i0 + 1 < n;
            const bool s0 = C[u].x == 0, s1 = has1 && C[u].y == 0;  // this phase's atoms
            double2 WX, WY, WZ;
            WX.x = WY.x = WZ.x = WX.y = WY.y = WZ.y = 0.0;
            if (s0) finish_atom(ss, zero, VX[u].x, VY[u].x, VZ[u].x, step, lambda, fc.hc, fc.mass, shift, WX.x, WY.x, WZ.x, nh);
            if (s1) finish_atom(ss, zero, VX[u].y, VY[u].y, VZ[u].y, step, lambda, fc.hc, fc.mass, shift, WX.y, WY.y, WZ.y, nh);
            if (s0 && s1) {
                if (store_state) {
                    const double2 z2 = make_double2(0.0, 0.0);
                    reinterpret_cast<double2 *>(a.fx)[t] = z2; reinterpret_cast<double2 *>(a.fy)[t] = z2;
                    reinterpret_cast<double2 *>(a.fz)[t] = z2; reinterpret_cast<double2 *>(a.u)[t] = z2;
                    reinterpret_cast<double2 *>(a.w)[t] = z2;
                    if (step) {
                        reinterpret_cast<double2 *>(a.vx)[t] = VX[u]; reinterpret_cast<double2 *>(a.vy)[t] = VY[u];
                        reinterpret_cast<double2 *>(a.vz)[t] = VZ[u];
                    }
                } else {
                    reinterpret_cast<double2 *>(a.vx)[t] = WX; reinterpret_cast<double2 *>(a.vy)[t] = WY;
                    reinterpret_cast<double2 *>(a.vz)[t] = WZ;
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (!(h ? s1 : s0)) continue;
                    const int i = i0 + h;
                    const double vx = h ? VX[u].y : VX[u].x, vy = h ? VY[u].y : VY[u].x, vz = h ? VZ[u].y : VZ[u].x;
                    const double wx = h ? WX.y : WX.x, wy = h ? WY.y : WY.x, wz = h ? WZ.y : WZ.x;
                    if (store_state) {
                        a.fx[i] = 0.0; a.fy[i] = 0.0; a.fz[i] = 0.0; a.u[i] = 0.0; a.w[i] = 0.0;
                        if (step) { a.vx[i] = vx; a.vy[i] = vy; a.vz[i] = vz; }
                    } else {
                        a.vx[i] = wx; a.vy[i] = wy; a.vz[i] = wz;
                    }
                }
            }
        }
    }

    // ---- phase A: atoms with partners, one per thread ---------------------------------------------------------------
    LjConst c;
    c.Lx = sc->box[0]; c.Ly = sc->box[1]; c.Lz = sc->box[2];
    c.hx = c.Lx / 2.0; c.hy = c.Ly / 2.0; c.hz = c.Lz / 2.0;
    c.hxi = __double2hiint(c.hx); c.hyi = __double2hiint(c.hy); c.hzi = __double2hiint(c.hz);
    const double *__restrict__ px = a.x, *__restrict__ py = a.y, *__restrict__ pz = a.z;
    const int n_active = *n_active_p;
    for (int k = blockIdx.x * FORCE_BLOCK + threadIdx.x; k < n_active; k += tstride) {
        const int i = active_idx[k];
        const double xi = px[i], yi = py[i], zi = pz[i];
        double vx = a.vx[i], vy = a.vy[i], vz = a.vz[i];
        const int cnt = nbr_cnt[i];
        int j = nbr[i];  // row 0
        PairAcc f = zero;
        for (int kk = 0; kk < cnt; ++kk) {
            const int jn = kk + 1 < cnt ? nbr[(size_t)(kk + 1) * npad + i] : 0;
            pair_fast_branchy(f, true, px[j], py[j], pz[j], xi, yi, zi, c, fc);
            j = jn;
        }
        double wx, wy, wz;
        finish_atom(ss, f, vx, vy, vz, step, lambda, fc.hc, fc.mass, shift, wx, wy, wz, nh);
        if (store_state) {
            a.fx[i] = f.fx; a.fy[i] = f.fy; a.fz[i] = f.fz; a.u[i] = f.u; a.w[i] = f.w;
            if (step) { a.vx[i] = vx; a.vy[i] = vy; a.vz[i] = vz; }
        } else {
            a.vx[i] = wx; a.vy[i] = wy; a.vz[i] = wz;
        }
    }

    Sums s;
#pragma unroll
    for (int q = 0; q < NSUM; ++q) s.v[q] = ss.v[q][threadIdx.x];
    block_reduce<FORCE_BLOCK>(s);
    grid_reduce_finalize<FORCE_BLOCK>(s, partials, sc, pr,
                                      (step ? FIN_STEP : 0) | (do_step & 2 ? FIN_DIST : 0) | (do_step & 8 ? FIN_P2P : 0),
                                      cond_handle, peers);
}

// Two consecutive atoms per thread: every plane access is one 128-bit transaction, all of a pair's loads are issued
// before the first use, and the neighbour loop advances both lists together (independent gather chains) with the
// next rows of partner indices prefetched while the current ones are in flight.
//   ROWS = 2: two list rows per trip (dense systems; 12 gathers in flight, 128 registers)
//   ROWS = 1: one row per trip (dilute systems: few partners, occupancy matters more than unrolling)
// UNION (dense only): nbr / nbr_cnt are the union table and its per-thread lengths (k_build_union), npad its row stride * 2
template <bool EXACT, int ROWS, bool MASKED, bool UNION = false>
__global__ void __launch_bounds__(FORCE_BLOCK, (EXACT || MASKED) ? MD_FORCE_MINB : MD_FORCE_MINB_DILUTE)
    k_force(int n, Arrays a, const int *__restrict__ nbr, const int *__restrict__ nbr_cnt, int npad, int cap,
            double *__restrict__ partials, Scalars *sc, const Params *__restrict__ pr, int do_step,
            unsigned long long cond_handle, const ForceConsts fc, const Peers *peers)
{
    // do_step bits: 1 = MD step (both half-kicks fused in), 2 = multi-GPU (publish rank sums only), 4 = guarded,
    //               8 = multi-GPU over peer memory (sums exchanged and finalized here), 16 = wait for the neighbours' ghosts,
    //               32 = launched as a programmatic dependent (wait for the predecessor first)
    if (!force_prologue(do_step, sc, peers)) return;
    if (threadIdx.x == 0) { PROBE_MIN(0); }
    __shared__ SumsSmem ss;
#pragma unroll
    for (int q = 0; q < NSUM; ++q) ss.v[q][threadIdx.x] = 0.0;
    // Control words are rewritten only by the last block's finalize, after every block has finished its atoms.
    const bool store_state = !(do_step & 1) || sc->steps_left <= 1;
    const bool nh = pr->th_kind == 2 || !(do_step & 1);  // a plain force evaluation keeps every stored sum valid
    const double lambda = MASKED ? 1.0 : sc->lambda;
    LjConst c;
    c.Lx = sc->box[0]; c.Ly = sc->box[1]; c.Lz = sc->box[2];
    c.hx = c.Lx / 2.0; c.hy = c.Ly / 2.0; c.hz = c.Lz / 2.0;
    c.hxi = __double2hiint(c.hx); c.hyi = __double2hiint(c.hy); c.hzi = __double2hiint(c.hz);
    const double shift[3] = {MASKED ? 0.0 : sc->shift[0], MASKED ? 0.0 : sc->shift[1], MASKED ? 0.0 : sc->shift[2]};
    // partners listed at the last build are now at most r_list + skin away (each atom moved < skin/2)
    const double wrap_margin = (2.0 * pr->r_list - pr->r_cut) * 1.02;
    const int npairs = (n + 1) >> 1;
    const int last_row = cap - 1;
    const double *__restrict__ px = a.x, *__restrict__ py = a.y, *__restrict__ pz = a.z;
    // Dilute / exact variants: the next pair's operands (x, y, z, u, list count, first list row: 112 B per thread) are
    // copied into shared memory by cp.async while the current pair's gathers and arithmetic run, so the streaming loads
    // overlap the latency-bound neighbour phase instead of alternating with it.  (The dense variant is bound by its
    // neighbour loop and keeps its L1 for gathers.)
    constexpr bool PREFETCH = !MASKED;
    __shared__ __align__(8) int2 ring_store[MASKED ? RING_D * 2 * FORCE_BLOCK : 1];
    IndexRing &ring = *reinterpret_cast<IndexRing *>(ring_store);
    // per-atom potential and virial enter nothing but the stored State and the S_U / S_W sums: the potential sum is only
    // reported, the virial sum feeds the barostat — steady-state steps of a batch skip what nobody reads
    const bool need_u = store_state, need_w = store_state || pr->ba_kind != 0;
    __shared__ __align__(16) double2 pf[PREFETCH ? 2 : 1][PREFETCH ? 6 : 1][PREFETCH ? FORCE_BLOCK : 1];
    __shared__ __align__(8) int2 pfi[PREFETCH ? 2 : 1][PREFETCH ? 2 : 1][PREFETCH ? FORCE_BLOCK : 1];
    const int tstride = gridDim.x * FORCE_BLOCK;
#define MD_PREFETCH_PAIR(S, TT)                                                                  \
    do {                                                                                         \
        const int l_ = threadIdx.x;                                                              \
        cp_async16(&pf[S][0][l_], reinterpret_cast<const double2 *>(px) + (TT));                 \
        cp_async16(&pf[S][1][l_], reinterpret_cast<const double2 *>(py) + (TT));                 \
        cp_async16(&pf[S][2][l_], reinterpret_cast<const double2 *>(pz) + (TT));                 \
        cp_async16(&pf[S][3][l_], reinterpret_cast<const double2 *>(a.vx) + (TT));               \
        cp_async16(&pf[S][4][l_], reinterpret_cast<const double2 *>(a.vy) + (TT));               \
        cp_async16(&pf[S][5][l_], reinterpret_cast<const double2 *>(a.vz) + (TT));               \
        cp_async8(&pfi[S][0][l_], reinterpret_cast<const int2 *>(nbr_cnt) + (TT));               \
        cp_async8(&pfi[S][1][l_], reinterpret_cast<const int2 *>(nbr) + (TT));                   \
        cp_async_commit();                                                                       \
    } while (0)
    int t = blockIdx.x * FORCE_BLOCK + threadIdx.x;
    if (PREFETCH && t < npairs) MD_PREFETCH_PAIR(0, t);
    for (int it = 0; t < npairs; t += tstride, ++it) {
        const int i0 = 2 * t;
        const bool has1 = i0 + 1 < n;
        const int2 *__restrict__ row = reinterpret_cast<const int2 *>(nbr) + t;
        const size_t stride = (size_t)(npad >> 1);
        double2 X, Y, Z, VX, VY, VZ;
        int2 C, J0;
        if (PREFETCH) {
            const int s = it & 1, l = threadIdx.x;
            cp_async_wait_all();
            X = pf[s][0][l]; Y = pf[s][1][l]; Z = pf[s][2][l];
            VX = pf[s][3][l]; VY = pf[s][4][l]; VZ = pf[s][5][l];
            C = pfi[s][0][l]; J0 = pfi[s][1][l];
            if (t + tstride < npairs) MD_PREFETCH_PAIR(s ^ 1, t + tstride);
        } else {
            X = reinterpret_cast<const double2 *>(px)[t]; Y = reinterpret_cast<const double2 *>(py)[t];
            Z = reinterpret_cast<const double2 *>(pz)[t];
            if (UNION) { C = make_int2(nbr_cnt[t], 0); J0 = make_int2(0, 0); }
            else { C = reinterpret_cast<const int2 *>(nbr_cnt)[t]; J0 = row[0]; }
            if (!MASKED) {  // dense: the velocities are fetched after the (long) neighbour loop — 12 registers less in it
                VX = reinterpret_cast<double2 *>(a.vx)[t]; VY = reinterpret_cast<double2 *>(a.vy)[t];
                VZ = reinterpret_cast<double2 *>(a.vz)[t];
            }
        }
        if (!has1) C.y = 0;
        PairAcc f0 = {0.0, 0.0, 0.0, 0.0, 0.0}, f1 = {0.0, 0.0, 0.0, 0.0, 0.0};
        if (EXACT) {
            for (int k = 0; k < C.x; ++k) {
                int j = k ? row[k * stride].x : J0.x;
                pair_exact(f0, px[j], py[j], pz[j], X.x, Y.x, Z.x, c, fc);
            }
            for (int k = 0; k < C.y; ++k) {
                int j = k ? row[k * stride].y : J0.y;
                pair_exact(f1, px[j], py[j], pz[j], X.y, Y.y, Z.y, c, fc);
            }
        } else {
            if (MASKED) {
                // warp-uniform choice: is any atom of this warp within r_list + skin of a box face?
                const double m = wrap_margin;
                const bool near = X.x < m || X.x > c.Lx - m || Y.x < m || Y.x > c.Ly - m || Z.x < m || Z.x > c.Lz - m ||
                                  X.y < m || X.y > c.Lx - m || Y.y < m || Y.y > c.Ly - m || Z.y < m || Z.y > c.Lz - m;
                const bool wrap = __any_sync(__activemask(), near);
                // uniform over the grid: one of six loop instances runs per launch
                const int uw = need_u ? 2 : (need_w ? 1 : 0);
                const int *urow = nbr + t;  // UNION: entry k of thread t is nbr[k * (npad / 2) + t]
#define MD_DENSE_CALL(W, U)                                                                                       \
    do {                                                                                                          \
        if (UNION) neighbour_loop_union<W, U>(f0, f1, a.q4, urow, stride, C.x, i0, X, Y, Z, c, fc, ring, n);      \
        else neighbour_loop_dense<W, U>(f0, f1, a.q4, row, stride, C, i0, X, Y, Z, c, fc, ring, n);               \
    } while (0)
                if (wrap) {
                    if (uw == 0) MD_DENSE_CALL(true, 0); else if (uw == 1) MD_DENSE_CALL(true, 1); else MD_DENSE_CALL(true, 2);
                } else {
                    if (uw == 0) MD_DENSE_CALL(false, 0); else if (uw == 1) MD_DENSE_CALL(false, 1); else MD_DENSE_CALL(false, 2);
                }
#undef MD_DENSE_CALL
            } else {
                neighbour_loop<ROWS, false, true>(f0, f1, a, row, stride, last_row, C, i0, X, Y, Z, c, fc, J0);
            }
        }
        if (MASKED) {
            VX = reinterpret_cast<double2 *>(a.vx)[t]; VY = reinterpret_cast<double2 *>(a.vy)[t];
            VZ = reinterpret_cast<double2 *>(a.vz)[t];
        }
        double2 WX, WY, WZ;
        // dense: lambda and the sum shift are re-read here rather than carried through the neighbour loop (7 registers)
        const double lam = MASKED ? ld_pinned(&sc->lambda) : lambda;
        const double sh[3] = {MASKED ? ld_pinned(&sc->shift[0]) : shift[0], MASKED ? ld_pinned(&sc->shift[1]) : shift[1],
                              MASKED ? ld_pinned(&sc->shift[2]) : shift[2]};
        finish_atom(ss, f0, VX.x, VY.x, VZ.x, (do_step & 1) != 0, lam, fc.hc, fc.mass, sh, WX.x, WY.x, WZ.x, nh);
        if (has1) finish_atom(ss, f1, VX.y, VY.y, VZ.y, (do_step & 1) != 0, lam, fc.hc, fc.mass, sh, WX.y, WY.y, WZ.y, nh);
        else { WX.y = WY.y = WZ.y = 0.0; }
        if (!has1) {  // odd tail: scalar stores only (slot i0+1 may hold a ghost atom in the distributed layout)
            if (store_state) {
                a.fx[i0] = f0.fx; a.fy[i0] = f0.fy; a.fz[i0] = f0.fz; a.u[i0] = f0.u; a.w[i0] = f0.w;
                if (do_step & 1) { a.vx[i0] = VX.x; a.vy[i0] = VY.x; a.vz[i0] = VZ.x; }
            } else {
                a.vx[i0] = WX.x; a.vy[i0] = WY.x; a.vz[i0] = WZ.x;
            }
        } else if (store_state) {
            reinterpret_cast<double2 *>(a.fx)[t] = make_double2(f0.fx, f1.fx);
            reinterpret_cast<double2 *>(a.fy)[t] = make_double2(f0.fy, f1.fy);
            reinterpret_cast<double2 *>(a.fz)[t] = make_double2(f0.fz, f1.fz);
            reinterpret_cast<double2 *>(a.u)[t] = make_double2(f0.u, f1.u);
            reinterpret_cast<double2 *>(a.w)[t] = make_double2(f0.w, f1.w);
            if (do_step & 1) {
                reinterpret_cast<double2 *>(a.vx)[t] = VX; reinterpret_cast<double2 *>(a.vy)[t] = VY;
                reinterpret_cast<double2 *>(a.vz)[t] = VZ;
            }
        } else {
            reinterpret_cast<double2 *>(a.vx)[t] = WX; reinterpret_cast<double2 *>(a.vy)[t] = WY;
            reinterpret_cast<double2 *>(a.vz)[t] = WZ;
        }
    }
    if (threadIdx.x == 0) { PROBE_MAX(1); }
    Sums s;
#pragma unroll
    for (int q = 0; q < NSUM; ++q) s.v[q] = ss.v[q][threadIdx.x];
    block_reduce<FORCE_BLOCK>(s);
    grid_reduce_finalize<FORCE_BLOCK>(s, partials, sc, pr,
                                      (do_step & 1 ? FIN_STEP : 0) | (do_step & 2 ? FIN_DIST : 0) | (do_step & 8 ? FIN_P2P : 0),
                                      cond_handle, peers);
}

// ----------------------------------------------------------------------------------------------------
// Dense systems, warp-cooperative variant (MD_FORCE_FAST_COOP).
//
// ncu on k_force<.., MASKED> (profiles/r01_ncu_c5_v10): 363 k L1 wavefronts per SM in 359 k active cycles — the per-thread
// Verlet loop is bound by the L1 tag stage, not by FP64 (31 % busy): lane l gathers partner k of ITS atom, 32 lanes hit 32
// different 128-byte lines, one wavefront each.  Here the 32 lanes of a warp work on ONE atom at a time: lane l takes list
// entries l, l + 32, ...  The list is stored atom-major (k_transpose_list), so the index read is one coalesced line, and
// since a list is the concatenation of ascending index runs (one per stencil cell run) consecutive entries are mostly
// consecutive atoms: four partners share a 128-byte line of the packed copy q4 and a gather costs ~8-12 wavefronts instead
// of 32.  The per-lane partial forces are folded with xor-shuffles (fixed order: deterministic) and handed to the lane
// that owns the atom, so everything after the neighbour phase is k_force's (two atoms per thread, 128-bit plane accesses).
// Index rows are staged COOP_STAGES atoms ahead in a per-warp shared-memory ring by cp.async (a lane reads back only what
// it copied: no barrier).
constexpr int COOP_ROWS = 7;     // rows of 32 entries gathered from registers per atom (224 partners); longer lists take the tail loop
constexpr int COOP_STAGES = 4;
#ifndef MD_COOP_MINB
#define MD_COOP_MINB 4
#endif

// nbr[k * npad + i] (k-major, one coalesced row per partner slot) -> nbrT[i * capT + k] (atom-major); 32 x 32 tiles
__global__ void __launch_bounds__(256) k_transpose_list(int n, int cap, int npad, int capT, const int *__restrict__ nbr,
                                                        int *__restrict__ nbrT)
{
    __shared__ int tile[32][33];
    const int i0 = blockIdx.x * 32, k0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        const int k = k0 + ty + r, i = i0 + tx;
        tile[ty + r][tx] = (k < cap && i < n) ? nbr[(size_t)k * npad + i] : 0;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 32; r += 8) {
        const int i = i0 + ty + r, k = k0 + tx;
        if (i < n && k < capT) nbrT[(size_t)i * capT + k] = tile[tx][ty + r];
    }
}

// R rows of one atom's list, straight-line: R gathers, then R pair terms.  (With the row count as a run-time condition
// ptxas sinks every gather into the branch that uses it, right in front of its pair term — the latencies then add up.)
template <bool WRAP, int UW, int R>
__device__ __forceinline__ void coop_rows(PairAcc &acc, const double4 *__restrict__ q4, const int *ring_stage, int cnt,
                                          int dummy, double xi, double yi, double zi, const LjConst &c,
                                          const ForceConsts &fc)
{
    const int lane = threadIdx.x & 31;
#define MD_ROW(M)                                                                                         \
    double2 p##M = make_double2(0.0, 0.0);                                                                \
    double z##M = 0.0;                                                                                    \
    if ((M) < R) {                                                                                        \
        const int k_ = (M) * 32 + lane;                                                                   \
        const int j_ = k_ < cnt ? ring_stage[k_] : dummy; /* entries past the list share one address */   \
        double w_;                                                                                        \
        asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"                                           \
                     : "=d"(p##M.x), "=d"(p##M.y), "=d"(z##M), "=d"(w_)                                   \
                     : "l"(q4 + j_));                                                                     \
    }
    MD_ROW(0) MD_ROW(1) MD_ROW(2) MD_ROW(3) MD_ROW(4) MD_ROW(5) MD_ROW(6)
#undef MD_ROW
#define MD_TERM(M) \
    if ((M) < R) pair_dense<WRAP, UW>(acc, (M) * 32 + lane < cnt, p##M.x, p##M.y, z##M, xi, yi, zi, c, fc);
    MD_TERM(0) MD_TERM(1) MD_TERM(2) MD_TERM(3) MD_TERM(4) MD_TERM(5) MD_TERM(6)
#undef MD_TERM
}

template <bool WRAP, int UW>
__device__ __forceinline__ void coop_atom(PairAcc &acc, const double4 *__restrict__ q4, const int *__restrict__ lst,
                                          const int *ring_stage, int cnt, int dummy, double xi, double yi, double zi,
                                          const LjConst &c, const ForceConsts &fc)
{
    static_assert(COOP_ROWS == 7, "coop_rows is written out for seven rows");
    const int rows = (cnt + 31) >> 5;  // warp-uniform
#define MD_CASE(R) case R: coop_rows<WRAP, UW, R>(acc, q4, ring_stage, cnt, dummy, xi, yi, zi, c, fc); break
    switch (min(rows, COOP_ROWS)) {
        MD_CASE(1); MD_CASE(2); MD_CASE(3); MD_CASE(4); MD_CASE(5); MD_CASE(6); MD_CASE(7);
        default: break;
    }
#undef MD_CASE
    const int lane = threadIdx.x & 31;
    for (int m = COOP_ROWS; m < rows; ++m) {  // lists beyond COOP_ROWS * 32 entries: straight from the table
        const int k = m * 32 + lane;
        const int j = k < cnt ? lst[k] : dummy;
        const double4 q = q4[j];
        pair_dense<WRAP, UW>(acc, k < cnt, q.x, q.y, q.z, xi, yi, zi, c, fc);
    }
}

__global__ void __launch_bounds__(FORCE_BLOCK, MD_COOP_MINB)
    k_force_coop(int n, Arrays a, const int *__restrict__ nbrT, const int *__restrict__ nbr_cnt, int capT,
                 double *__restrict__ partials, Scalars *sc, const Params *__restrict__ pr, int do_step,
                 unsigned long long cond_handle, const ForceConsts fc, const Peers *peers)
{
    if (!force_prologue(do_step, sc, peers)) return;
    __shared__ SumsSmem ss;
#pragma unroll
    for (int q = 0; q < NSUM; ++q) ss.v[q][threadIdx.x] = 0.0;
    __shared__ int ring_store[FORCE_BLOCK / 32][COOP_STAGES][COOP_ROWS * 32];
    const bool store_state = !(do_step & 1) || sc->steps_left <= 1;
    const bool nh = pr->th_kind == 2 || !(do_step & 1);
    const bool need_u = store_state, need_w = store_state || pr->ba_kind != 0;
    const int uw = need_u ? 2 : (need_w ? 1 : 0);
    LjConst c;
    c.Lx = sc->box[0]; c.Ly = sc->box[1]; c.Lz = sc->box[2];
    c.hx = c.Lx / 2.0; c.hy = c.Ly / 2.0; c.hz = c.Lz / 2.0;
    c.hxi = c.hyi = c.hzi = 0;
    const double wrap_margin = (2.0 * pr->r_list - pr->r_cut) * 1.02;  // see k_force
    const int npairs = (n + 1) >> 1;
    const int tstride = gridDim.x * FORCE_BLOCK;
    const int lane = threadIdx.x & 31;
    int(*ring)[COOP_ROWS * 32] = ring_store[threadIdx.x >> 5];
    for (int tb = blockIdx.x * FORCE_BLOCK + threadIdx.x - lane; tb < npairs; tb += tstride) {  // warp-uniform
        const int t = tb + lane;
        const bool valid = t < npairs;
        const int i0 = 2 * t;
        const bool has1 = valid && i0 + 1 < n;
        int2 C = valid ? reinterpret_cast<const int2 *>(nbr_cnt)[t] : make_int2(0, 0);
        if (!has1) C.y = 0;
        const int base = 2 * tb;                       // first atom of the warp's tile (multiple of 64)
        const int na = min(64, n - base);              // atoms in the tile
        const int dummy = safe_dummy(base, n);         // never one of the tile's atoms (n >= 128 on this path)
        PairAcc f0 = {0.0, 0.0, 0.0, 0.0, 0.0}, f1 = {0.0, 0.0, 0.0, 0.0, 0.0};
        // stage the index rows of atom A (a lane copies the entries it will read itself); one commit group per atom
#define MD_STAGE_ATOM(A)                                                                                          \
    do {                                                                                                          \
        const int a_ = (A);                                                                                       \
        if (a_ < na) {                                                                                            \
            const int cnt_ = __shfl_sync(0xffffffffu, (a_ & 1) ? C.y : C.x, a_ >> 1);                             \
            const int *src_ = nbrT + (size_t)(base + a_) * capT;                                                  \
            int *dst_ = ring[a_ % COOP_STAGES];                                                                   \
            _Pragma("unroll") for (int m = 0; m < COOP_ROWS; ++m) {                                               \
                const int k_ = m * 32 + lane;                                                                     \
                if (k_ < cnt_)                                                                                    \
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst_ + k_)), "l"(src_ + k_) \
                                 : "memory");                                                                     \
            }                                                                                                     \
        }                                                                                                         \
        cp_async_commit();                                                                                        \
    } while (0)
#pragma unroll
        for (int s = 0; s < COOP_STAGES - 1; ++s) MD_STAGE_ATOM(s);
        for (int at = 0; at < na; ++at) {
            MD_STAGE_ATOM(at + COOP_STAGES - 1);
            asm volatile("cp.async.wait_group %0;" ::"n"(COOP_STAGES - 1) : "memory");
            const int cnt = __shfl_sync(0xffffffffu, (at & 1) ? C.y : C.x, at >> 1);
            const int i = base + at;
            const double4 qi = a.q4[i];  // warp-uniform address
            const double m = wrap_margin;
            const bool wrap = qi.x < m || qi.x > c.Lx - m || qi.y < m || qi.y > c.Ly - m || qi.z < m || qi.z > c.Lz - m;
            PairAcc acc = {0.0, 0.0, 0.0, 0.0, 0.0};
            const int *lst = nbrT + (size_t)i * capT;
            const int *stage = ring[at % COOP_STAGES];
#define MD_COOP_CALL(W, U) coop_atom<W, U>(acc, a.q4, lst, stage, cnt, dummy, qi.x, qi.y, qi.z, c, fc)
            if (wrap) {
                if (uw == 0) MD_COOP_CALL(true, 0); else if (uw == 1) MD_COOP_CALL(true, 1); else MD_COOP_CALL(true, 2);
            } else {
                if (uw == 0) MD_COOP_CALL(false, 0); else if (uw == 1) MD_COOP_CALL(false, 1); else MD_COOP_CALL(false, 2);
            }
#undef MD_COOP_CALL
            // fixed-order fold over the lanes
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc.fx += __shfl_xor_sync(0xffffffffu, acc.fx, o);
                acc.fy += __shfl_xor_sync(0xffffffffu, acc.fy, o);
                acc.fz += __shfl_xor_sync(0xffffffffu, acc.fz, o);
                if (uw >= 1) acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
                if (uw >= 2) acc.u += __shfl_xor_sync(0xffffffffu, acc.u, o);
            }
            if (lane == (at >> 1)) {
                if (at & 1) f1 = acc; else f0 = acc;
            }
        }
#undef MD_STAGE_ATOM
        cp_async_wait_all();
        if (!valid) continue;
        // ---- from here on: k_force's epilogue for the two atoms this thread owns -------------------------------------
        double2 VX = reinterpret_cast<double2 *>(a.vx)[t], VY = reinterpret_cast<double2 *>(a.vy)[t],
                VZ = reinterpret_cast<double2 *>(a.vz)[t];
        double2 WX, WY, WZ;
        const double lam = ld_pinned(&sc->lambda);
        const double sh[3] = {ld_pinned(&sc->shift[0]), ld_pinned(&sc->shift[1]), ld_pinned(&sc->shift[2])};
        finish_atom(ss, f0, VX.x, VY.x, VZ.x, (do_step & 1) != 0, lam, fc.hc, fc.mass, sh, WX.x, WY.x, WZ.x, nh);
        if (has1) finish_atom(ss, f1, VX.y, VY.y, VZ.y, (do_step & 1) != 0, lam, fc.hc, fc.mass, sh, WX.y, WY.y, WZ.y, nh);
        else { WX.y = WY.y = WZ.y = 0.0; }
        if (!has1) {
            if (store_state) {
                a.fx[i0] = f0.fx; a.fy[i0] = f0.fy; a.fz[i0] = f0.fz; a.u[i0] = f0.u; a.w[i0] = f0.w;
                if (do_step & 1) { a.vx[i0] = VX.x; a.vy[i0] = VY.x; a.vz[i0] = VZ.x; }
            } else {
                a.vx[i0] = WX.x; a.vy[i0] = WY.x; a.vz[i0] = WZ.x;
            }
        } else if (store_state) {
            reinterpret_cast<double2 *>(a.fx)[t] = make_double2(f0.fx, f1.fx);
            reinterpret_cast<double2 *>(a.fy)[t] = make_double2(f0.fy, f1.fy);
            reinterpret_cast<double2 *>(a.fz)[t] = make_double2(f0.fz, f1.fz);
            reinterpret_cast<double2 *>(a.u)[t] = make_double2(f0.u, f1.u);
            reinterpret_cast<double2 *>(a.w)[t] = make_double2(f0.w, f1.w);
            if (do_step & 1) {
                reinterpret_cast<double2 *>(a.vx)[t] = VX; reinterpret_cast<double2 *>(a.vy)[t] = VY;
                reinterpret_cast<double2 *>(a.vz)[t] = VZ;
            }
        } else {
            reinterpret_cast<double2 *>(a.vx)[t] = WX; reinterpret_cast<double2 *>(a.vy)[t] = WY;
            reinterpret_cast<double2 *>(a.vz)[t] = WZ;
        }
    }
    Sums s;
#pragma unroll
    for (int q = 0; q < NSUM; ++q) s.v[q] = ss.v[q][threadIdx.x];
    block_reduce<FORCE_BLOCK>(s);
    grid_reduce_finalize<FORCE_BLOCK>(s, partials, sc, pr,
                                      (do_step & 1 ? FIN_STEP : 0) | (do_step & 2 ? FIN_DIST : 0) | (do_step & 8 ? FIN_P2P : 0),
                                      cond_handle, peers);
}

// ----------------------------------------------------------------------------------------------------
// K4: (first half-kick,) thermostat scale, pending barostat coordinate scale, drift, periodic wrap.
//   integrator.rs:28-34  u = v + F*(dt/(2m))    only on the first step of a batch; afterwards k_force left u
//   thermostat.rs:54-58  v' = u*lambda          (lambda == 1.0 without thermostat: bitwise no-op)
//   barostat.rs:46-48    x *= myu of the previous step (mu_pending == 1.0 otherwise: bitwise no-op)
//   integrator.rs:40-44  x += v'*dt
//   particle.rs:120-142  single-shift wrap into [0, L)
// Element-wise and HBM-bound: two atoms per thread, 128-bit accesses; explicit _rn intrinsics keep the
// reference's rounding (no FMA contraction).  v' itself is not stored: k_force recomputes the same product.
__device__ __forceinline__ void drift_one(double &x, double u, double lambda, double mup, double dt, double L)
{
    double v = __dmul_rn(u, lambda);
    x = __dmul_rn(x, mup);
    x = __dadd_rn(x, __dmul_rn(v, dt));
    if (x < 0.0) x = __dadd_rn(x, L);
    else if (x >= L) x = __dsub_rn(x, L);
}

// L2 (PDL = true: the kernel may run while its predecessor is finishing, never trust an L1 line) or plain loads
template <bool PDL, typename T>
__device__ __forceinline__ T ld_state(const T *p)
{
    if constexpr (PDL) return __ldcg(p);
    else return *p;
}

template <bool PDL>
__device__ __forceinline__ void kick_drift_tail(int i, Arrays a, double lambda, double mup, double Lx, double Ly, double Lz,
                                                bool half, const Params *__restrict__ pr, bool write_q4)
{
    const double c = pr->half_dt_m, dt = pr->dt;
    double ux = ld_state<PDL>(a.vx + i), uy = ld_state<PDL>(a.vy + i), uz = ld_state<PDL>(a.vz + i);
    if (!half) {
        ux = __dadd_rn(ux, __dmul_rn(ld_state<PDL>(a.fx + i), c)); uy = __dadd_rn(uy, __dmul_rn(ld_state<PDL>(a.fy + i), c));
        uz = __dadd_rn(uz, __dmul_rn(ld_state<PDL>(a.fz + i), c));
        a.vx[i] = ux; a.vy[i] = uy; a.vz[i] = uz;
    }
    double x = ld_state<PDL>(a.x + i), y = ld_state<PDL>(a.y + i), z = ld_state<PDL>(a.z + i);
    drift_one(x, ux, lambda, mup, dt, Lx);
    drift_one(y, uy, lambda, mup, dt, Ly);
    drift_one(z, uz, lambda, mup, dt, Lz);
    a.x[i] = x; a.y[i] = y; a.z[i] = z;
    if (write_q4) a.q4[i] = make_double4(x, y, z, 0.0);
}

// Multi-GPU over peer memory: the face atoms of a slab are a prefix [0, m_left) and a suffix [n - m_right, n) of its
// cell-sorted order (ghosts are selected by x cell layer), so the drift kernel itself stores their new positions into the
// neighbours' ghost slots — plain NVLink stores into the neighbour's HBM, no fence here.  The kernel boundary orders them;
// the first thing k_force does is raise the step's sequence flag in both neighbours' mailboxes and poll its own.
struct HaloPush {
    int m[2];                    // face atoms for the left / right neighbour (0, 0: nothing to push, e.g. single GPU)
    double *x[2], *y[2], *z[2];  // the neighbour's planes (mapped), already offset to the first ghost slot we own there
    double4 *q4[2];
};

__device__ __forceinline__ void push_atom(const HaloPush &h, int i, int n, double x, double y, double z)
{
    if (i < h.m[0]) {
        h.x[0][i] = x; h.y[0][i] = y; h.z[0][i] = z;
        if (h.q4[0]) h.q4[0][i] = make_double4(x, y, z, 0.0);
    }
    const int k = i - (n - h.m[1]);
    if (k >= 0) {
        h.x[1][k] = x; h.y[1][k] = y; h.z[1][k] = z;
        if (h.q4[1]) h.q4[1][k] = make_double4(x, y, z, 0.0);
    }
}

// PDL = false: plain launch, the predecessor is complete (every launch but the ones below).  PDL = true: launched as a
// programmatic dependent of k_force inside a single-GPU chunk graph (MOLDYN_B200_PDL, opt-in).
template <bool PDL>
__global__ void __launch_bounds__(256, 4) k_kick_drift(int n, Arrays a, Scalars *sc,
                                                    const Params *__restrict__ pr, int guarded, int write_q4,
                                                    int early_k, unsigned force_grid, const HaloPush h)
{
    // guarded bits: 1 = return at once when the loop is halted, 2 = programmatic dependent (== PDL),
    //               4 = the chunk's drifts start early (with 2: step early_k >= 1 of the chunk; step 0 is a plain launch)
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    double2 x, y, z, ux, uy, uz;
    bool have_x = false, have_u = false;
    double lambda, mup, Lx, Ly, Lz;
    bool half;
    if constexpr (PDL) {
        pdl_launch_dependents();  // k_force of this step may become resident; it waits for this grid to complete
        // positions were last written by the previous k_kick_drift, which completed before our predecessor (k_force) did
        // anything: they can be fetched while k_force drains.
        if (2 * t + 1 < n) {
            x = __ldcg(reinterpret_cast<const double2 *>(a.x) + t); y = __ldcg(reinterpret_cast<const double2 *>(a.y) + t);
            z = __ldcg(reinterpret_cast<const double2 *>(a.z) + t);
            have_x = true;
        }
        bool waited = false;
        if (guarded & 4) {
            // Early start.  The predecessor's tail — one block folding 592 partial sums and computing lambda, myu and the
            // rebuild decision while 147 SMs idle — is hidden behind this kernel's loads: (a) once every block of k_force has
            // taken its ticket, all velocities u' are final (each block fences before the ticket): fetch them; (b) once the
            // last block has release-stored the sequence number of this step, the controls are final: drift and store.
            // One thread per block polls (bounded); on a timeout, or on anything unexpected, the block falls back to
            // griddepcontrol.wait — always correct, the flags only ever let it start sooner.
            __shared__ int verdict;  // 0 = go, 1 = fall back to the full wait, 2 = halted: nothing to do
            __shared__ unsigned long long expect_s;
            if (threadIdx.x == 0) {
                const unsigned long long expect = __ldcg(&sc->chunk_fin0) + (unsigned long long)early_k;
                expect_s = expect;
                int v = 1;
                if (halted_now(sc)) v = 2;  // halted before our predecessor started: it is a no-op and raises no flag
                else
                    for (int spin = 0; spin < 4096; ++spin) {
                        if (ld_acquire_gpu(&sc->fin_seq) >= expect) { v = 3; break; }  // the whole predecessor is done
                        if (ld_acquire_gpu(&sc->ticket) == force_grid) { v = 0; break; }
                        __nanosleep(64);
                    }
                verdict = v;
            }
            __syncthreads();
            if (verdict == 2) return;
            if (verdict == 3) waited = true;
            else if (verdict == 0) {
                if (have_x) {
                    ux = __ldcg(reinterpret_cast<const double2 *>(a.vx) + t); uy = __ldcg(reinterpret_cast<const double2 *>(a.vy) + t);
                    uz = __ldcg(reinterpret_cast<const double2 *>(a.vz) + t);
                    have_u = true;
                }
                __syncthreads();  // verdict is rewritten below
                if (threadIdx.x == 0) {
                    const unsigned long long expect = expect_s;
                    int v = 1;
                    for (int spin = 0; spin < 4096; ++spin) {
                        if (ld_acquire_gpu(&sc->fin_seq) >= expect) { v = 0; break; }
                        __nanosleep(64);
                    }
                    verdict = v;
                }
                __syncthreads();
                waited = verdict == 0;
            }
        }
        if (!waited) pdl_wait();
        // The step controls, once per block through L2 (never a stale L1 line, and not 2000 blocks x 8 warps hammering one
        // L2 slice with the same nine words: measured 13 -> 26 us per launch at 10^6 atoms when every thread read them itself).
        __shared__ double ctl[5];   // lambda, mu_pending, Lx, Ly, Lz
        __shared__ int ctl_half, ctl_halted;
        if (threadIdx.x == 0) {
            const double l0 = __ldcg(&sc->lambda), l1 = __ldcg(&sc->mu_pending), l2 = __ldcg(&sc->box[0]),
                         l3 = __ldcg(&sc->box[1]), l4 = __ldcg(&sc->box[2]);
            ctl_half = __ldcg(&sc->vel_is_half);
            ctl_halted = halted_now(sc) ? 1 : 0;
            ctl[0] = l0; ctl[1] = l1; ctl[2] = l2; ctl[3] = l3; ctl[4] = l4;
        }
        __syncthreads();
        if ((guarded & 1) && ctl_halted) return;
        lambda = ctl[0]; mup = ctl[1]; Lx = ctl[2]; Ly = ctl[3]; Lz = ctl[4];
        half = ctl_half != 0;
    } else {
        if ((guarded & 1) && halted(sc)) return;
        lambda = sc->lambda; mup = sc->mu_pending;
        Lx = sc->box[0]; Ly = sc->box[1]; Lz = sc->box[2];
        half = sc->vel_is_half != 0;
        // first step of a chunk whose later drifts start early: the sequence number the chunk counts from
        if ((guarded & 4) && early_k == 0 && t == 0) sc->chunk_fin0 = sc->fin_seq;
    }
    // block-uniform: does this block hold face atoms?  (512 atoms per block)
    const int b_lo = blockIdx.x * 512, b_hi = b_lo + 512;
    const bool pushes = (h.m[0] | h.m[1]) != 0 && (b_lo < h.m[0] || b_hi > n - h.m[1]);
    if (2 * t < n) {
        if (2 * t + 1 >= n) {  // odd tail: one atom, scalar accesses (the slot after it may belong to a ghost atom)
            kick_drift_tail<PDL>(2 * t, a, lambda, mup, Lx, Ly, Lz, half, pr, write_q4 != 0);
            if (pushes) push_atom(h, 2 * t, n, a.x[2 * t], a.y[2 * t], a.z[2 * t]);
        } else {
            const double c = pr->half_dt_m, dt = pr->dt;
            if (!have_x) {
                x = reinterpret_cast<double2 *>(a.x)[t]; y = reinterpret_cast<double2 *>(a.y)[t];
                z = reinterpret_cast<double2 *>(a.z)[t];
            }
            if (!have_u) {
                ux = reinterpret_cast<double2 *>(a.vx)[t]; uy = reinterpret_cast<double2 *>(a.vy)[t];
                uz = reinterpret_cast<double2 *>(a.vz)[t];
            }
            if (!half) {
                const double2 fx = ld_state<PDL>(reinterpret_cast<const double2 *>(a.fx) + t),
                              fy = ld_state<PDL>(reinterpret_cast<const double2 *>(a.fy) + t),
                              fz = ld_state<PDL>(reinterpret_cast<const double2 *>(a.fz) + t);
                ux.x = __dadd_rn(ux.x, __dmul_rn(fx.x, c)); ux.y = __dadd_rn(ux.y, __dmul_rn(fx.y, c));
                uy.x = __dadd_rn(uy.x, __dmul_rn(fy.x, c)); uy.y = __dadd_rn(uy.y, __dmul_rn(fy.y, c));
                uz.x = __dadd_rn(uz.x, __dmul_rn(fz.x, c)); uz.y = __dadd_rn(uz.y, __dmul_rn(fz.y, c));
                reinterpret_cast<double2 *>(a.vx)[t] = ux; reinterpret_cast<double2 *>(a.vy)[t] = uy;
                reinterpret_cast<double2 *>(a.vz)[t] = uz;
            }
            drift_one(x.x, ux.x, lambda, mup, dt, Lx); drift_one(x.y, ux.y, lambda, mup, dt, Lx);
            drift_one(y.x, uy.x, lambda, mup, dt, Ly); drift_one(y.y, uy.y, lambda, mup, dt, Ly);
            drift_one(z.x, uz.x, lambda, mup, dt, Lz); drift_one(z.y, uz.y, lambda, mup, dt, Lz);
            reinterpret_cast<double2 *>(a.x)[t] = x; reinterpret_cast<double2 *>(a.y)[t] = y;
            reinterpret_cast<double2 *>(a.z)[t] = z;
            if (write_q4) {
                a.q4[2 * t] = make_double4(x.x, y.x, z.x, 0.0);
                a.q4[2 * t + 1] = make_double4(x.y, y.y, z.y, 0.0);
            }
            if (pushes) {
                push_atom(h, 2 * t, n, x.x, y.x, z.x);
                push_atom(h, 2 * t + 1, n, x.y, y.y, z.y);
            }
        }
    }
}

// ----------------------------------------------------------------------------------------------------
// K3+K4 fused — ONE kernel per step for dilute systems (few listed partners per atom).
//
// k_kick_drift exists as a separate kernel only because the forces need every partner's drifted position.  A thread
// can just as well drift its partners itself: x_j' = wrap(x_j*mu + (u_j*lambda)*dt) is the same instruction sequence
// the owner of j runs, hence the same bits.  With ~0.5 partners per atom that costs a few extra gathers and saves a
// full pass over the state: the step reads x,u (48 B/atom) + list count and first row (8 B) and writes x',u' (48 B).
// In-place updates would race with those partner reads, so x and v ping-pong between two plane sets (sc->parity
// names the current one; the last block flips it).
//
// Streaming side: each block walks its tiles of STEP_TILE atoms; the tile's eight plane segments are fetched by TMA
// bulk copies (cp.async.bulk → shared memory, mbarrier completion) into a two-stage ring, so the next tile's HBM
// requests are in flight while the block is busy with gathers and arithmetic of the current tile.
constexpr int STEP_TILE = 2 * FORCE_BLOCK;

struct StepStage {
    double x[STEP_TILE], y[STEP_TILE], z[STEP_TILE], ux[STEP_TILE], uy[STEP_TILE], uz[STEP_TILE];
    int cnt[STEP_TILE], row0[STEP_TILE];
};

__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// 1-D TMA bulk copy global → shared; bytes and both addresses are multiples of 16
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, unsigned bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// drifted position of an atom from its stored (x, u): thermostat.rs:54-58, barostat.rs:46-48, integrator.rs:40-45
__device__ __forceinline__ void drift3(double &x, double &y, double &z, double ux, double uy, double uz, double lambda,
                                       double mup, double dt, const LjConst &c)
{
    drift_one(x, ux, lambda, mup, dt, c.Lx);
    drift_one(y, uy, lambda, mup, dt, c.Ly);
    drift_one(z, uz, lambda, mup, dt, c.Lz);
}

#ifndef MD_STEP_MINB
#define MD_STEP_MINB 4
#endif
template <bool EXACT>
__global__ void __launch_bounds__(FORCE_BLOCK, MD_STEP_MINB)
    k_step_dilute(int n, Arrays P0, Arrays P1, const int *__restrict__ nbr, const int *__restrict__ nbr_cnt, int npad,
                  int cap, double *__restrict__ partials, Scalars *sc, const Params *__restrict__ pr, int flags,
                  unsigned long long cond_handle, const ForceConsts fc)
{
    if ((flags & 4) && halted(sc)) return;  // uniform over the grid: nobody takes a ticket
    __shared__ __align__(128) StepStage stg[2];
    __shared__ SumsSmem ss;
    __shared__ __align__(8) unsigned long long full[2];
    const int tid = threadIdx.x;
#pragma unroll
    for (int q = 0; q < NSUM; ++q) ss.v[q][tid] = 0.0;
    // Control words are rewritten only by the last block's finalize, after every block has finished its atoms.
    const bool par = sc->parity != 0;
    const bool store_state = sc->steps_left <= 1;
    const bool nh = pr->th_kind == 2;
    const double lambda = sc->lambda, mup = sc->mu_pending, dt = pr->dt;
    LjConst c;
    c.Lx = sc->box[0]; c.Ly = sc->box[1]; c.Lz = sc->box[2];
    c.hx = c.Lx / 2.0; c.hy = c.Ly / 2.0; c.hz = c.Lz / 2.0;
    c.hxi = __double2hiint(c.hx); c.hyi = __double2hiint(c.hy); c.hzi = __double2hiint(c.hz);
    const double shift[3] = {sc->shift[0], sc->shift[1], sc->shift[2]};
    const double *__restrict__ ix = par ? P1.x : P0.x, *__restrict__ iy = par ? P1.y : P0.y,
                 *__restrict__ iz = par ? P1.z : P0.z, *__restrict__ iux = par ? P1.vx : P0.vx,
                 *__restrict__ iuy = par ? P1.vy : P0.vy, *__restrict__ iuz = par ? P1.vz : P0.vz;
    double *__restrict__ ox = par ? P0.x : P1.x, *__restrict__ oy = par ? P0.y : P1.y, *__restrict__ oz = par ? P0.z : P1.z,
           *__restrict__ ovx = par ? P0.vx : P1.vx, *__restrict__ ovy = par ? P0.vy : P1.vy,
           *__restrict__ ovz = par ? P0.vz : P1.vz;
    const int ntiles = (n + STEP_TILE - 1) / STEP_TILE;
    const size_t stride = (size_t)(npad >> 1);

    auto issue = [&](int tile, int s) {  // one thread: arm the barrier, launch the eight segment copies
        const int base = tile * STEP_TILE;
        const unsigned na = (unsigned)min(STEP_TILE, npad - base);  // npad is a multiple of 64 atoms
        mbar_expect_tx(&full[s], na * 56u);
        tma_load_1d(stg[s].x, ix + base, na * 8u, &full[s]);
        tma_load_1d(stg[s].y, iy + base, na * 8u, &full[s]);
        tma_load_1d(stg[s].z, iz + base, na * 8u, &full[s]);
        tma_load_1d(stg[s].ux, iux + base, na * 8u, &full[s]);
        tma_load_1d(stg[s].uy, iuy + base, na * 8u, &full[s]);
        tma_load_1d(stg[s].uz, iuz + base, na * 8u, &full[s]);
        tma_load_1d(stg[s].cnt, nbr_cnt + base, na * 4u, &full[s]);
        tma_load_1d(stg[s].row0, nbr + base, na * 4u, &full[s]);
    };
    if (tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        if ((int)blockIdx.x < ntiles) issue(blockIdx.x, 0);
        if ((int)(blockIdx.x + gridDim.x) < ntiles) issue(blockIdx.x + gridDim.x, 1);
    }
    __syncthreads();

    for (int it = 0;; ++it) {
        const int tile = blockIdx.x + it * gridDim.x;
        if (tile >= ntiles) break;
        const int s = it & 1;
        mbar_wait(&full[s], (unsigned)(it >> 1) & 1u);
        double2 X = reinterpret_cast<const double2 *>(stg[s].x)[tid], Y = reinterpret_cast<const double2 *>(stg[s].y)[tid],
                Z = reinterpret_cast<const double2 *>(stg[s].z)[tid];
        double2 VX = reinterpret_cast<const double2 *>(stg[s].ux)[tid], VY = reinterpret_cast<const double2 *>(stg[s].uy)[tid],
                VZ = reinterpret_cast<const double2 *>(stg[s].uz)[tid];
        int2 C = reinterpret_cast<const int2 *>(stg[s].cnt)[tid];
        int2 J = reinterpret_cast<const int2 *>(stg[s].row0)[tid];
        const int i0 = tile * STEP_TILE + 2 * tid;
        const bool has0 = i0 < n, has1 = i0 + 1 < n;
        if (!has0) C.x = 0;
        if (!has1) C.y = 0;
        // own atoms: thermostat scale, pending barostat scale, drift, wrap
        drift3(X.x, Y.x, Z.x, VX.x, VY.x, VZ.x, lambda, mup, dt, c);
        drift3(X.y, Y.y, Z.y, VX.y, VY.y, VZ.y, lambda, mup, dt, c);
        PairAcc f0 = {0.0, 0.0, 0.0, 0.0, 0.0}, f1 = {0.0, 0.0, 0.0, 0.0, 0.0};
        const int kmax = max(C.x, C.y);
        const int2 *__restrict__ row = reinterpret_cast<const int2 *>(nbr) + (size_t)(i0 >> 1);
        for (int k = 0; k < kmax; ++k) {
            const bool a0 = k < C.x, a1 = k < C.y;
            const int j0 = a0 ? J.x : 0, j1 = a1 ? J.y : 0;
            if (k + 1 < kmax) J = row[(size_t)(k + 1) * stride];
            // all twelve gathers of this trip are issued before the first use
            double xa = __ldg(ix + j0), ya = __ldg(iy + j0), za = __ldg(iz + j0);
            const double uxa = __ldg(iux + j0), uya = __ldg(iuy + j0), uza = __ldg(iuz + j0);
            double xb = __ldg(ix + j1), yb = __ldg(iy + j1), zb = __ldg(iz + j1);
            const double uxb = __ldg(iux + j1), uyb = __ldg(iuy + j1), uzb = __ldg(iuz + j1);
            if (a0) {
                drift3(xa, ya, za, uxa, uya, uza, lambda, mup, dt, c);
                if (EXACT) pair_exact(f0, xa, ya, za, X.x, Y.x, Z.x, c, fc);
                else pair_fast_branchy(f0, true, xa, ya, za, X.x, Y.x, Z.x, c, fc);
            }
            if (a1) {
                drift3(xb, yb, zb, uxb, uyb, uzb, lambda, mup, dt, c);
                if (EXACT) pair_exact(f1, xb, yb, zb, X.y, Y.y, Z.y, c, fc);
                else pair_fast_branchy(f1, true, xb, yb, zb, X.y, Y.y, Z.y, c, fc);
            }
        }
        double2 WX, WY, WZ;
        WX.x = WY.x = WZ.x = WX.y = WY.y = WZ.y = 0.0;
        if (has0) finish_atom(ss, f0, VX.x, VY.x, VZ.x, true, lambda, fc.hc, fc.mass, shift, WX.x, WY.x, WZ.x, nh);
        if (has1) finish_atom(ss, f1, VX.y, VY.y, VZ.y, true, lambda, fc.hc, fc.mass, shift, WX.y, WY.y, WZ.y, nh);
        if (has1) {
            const int t = i0 >> 1;
            reinterpret_cast<double2 *>(ox)[t] = X; reinterpret_cast<double2 *>(oy)[t] = Y;
            reinterpret_cast<double2 *>(oz)[t] = Z;
            if (store_state) {
                reinterpret_cast<double2 *>(ovx)[t] = VX; reinterpret_cast<double2 *>(ovy)[t] = VY;
                reinterpret_cast<double2 *>(ovz)[t] = VZ;
                reinterpret_cast<double2 *>(P0.fx)[t] = make_double2(f0.fx, f1.fx);
                reinterpret_cast<double2 *>(P0.fy)[t] = make_double2(f0.fy, f1.fy);
                reinterpret_cast<double2 *>(P0.fz)[t] = make_double2(f0.fz, f1.fz);
                reinterpret_cast<double2 *>(P0.u)[t] = make_double2(f0.u, f1.u);
                reinterpret_cast<double2 *>(P0.w)[t] = make_double2(f0.w, f1.w);
            } else {
                reinterpret_cast<double2 *>(ovx)[t] = WX; reinterpret_cast<double2 *>(ovy)[t] = WY;
                reinterpret_cast<double2 *>(ovz)[t] = WZ;
            }
        } else if (has0) {  // odd tail: scalar stores only
            ox[i0] = X.x; oy[i0] = Y.x; oz[i0] = Z.x;
            if (store_state) {
                ovx[i0] = VX.x; ovy[i0] = VY.x; ovz[i0] = VZ.x;
                P0.fx[i0] = f0.fx; P0.fy[i0] = f0.fy; P0.fz[i0] = f0.fz; P0.u[i0] = f0.u; P0.w[i0] = f0.w;
            } else {
                ovx[i0] = WX.x; ovy[i0] = WY.x; ovz[i0] = WZ.x;
            }
        }
        // Refill stage s for the tile after next.  The TMA engine writes shared memory through the async proxy, which
        // is not ordered against shared-memory loads that are merely *issued*: the barrier therefore sits at the END of
        // the iteration, where every thread has consumed (stored results computed from) what it loaded from the stage.
        // (With the barrier right after the loads, a backed-up LSU queue let the refill overtake a warp's LDS.)
        __syncthreads();
        if (tid == 0) {
            const int next = tile + 2 * gridDim.x;
            if (next < ntiles) issue(next, s);
        }
    }
    Sums sum;
#pragma unroll
    for (int q = 0; q < NSUM; ++q) sum.v[q] = ss.v[q][tid];
    block_reduce<FORCE_BLOCK>(sum);
    grid_reduce_finalize<FORCE_BLOCK>(sum, partials, sc, pr, FIN_STEP | FIN_FLIP | (flags & 2 ? FIN_DIST : 0), cond_handle,
                                      nullptr);
}

// First step of a batch for the fused path: the velocity planes hold v (not u = v + F c) after an upload or after the
// last step of the previous batch (integrator.rs:28-34).
__global__ void k_first_half_kick(int n, Arrays a, Scalars *sc, const Params *__restrict__ pr)
{
    if (sc->vel_is_half) return;  // rewritten only by k_mark_half, a separate launch
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double c = pr->half_dt_m;
    a.vx[i] = __dadd_rn(a.vx[i], __dmul_rn(a.fx[i], c));
    a.vy[i] = __dadd_rn(a.vy[i], __dmul_rn(a.fy[i], c));
    a.vz[i] = __dadd_rn(a.vz[i], __dmul_rn(a.fz[i], c));
}

__global__ void k_mark_half(Scalars *sc) { sc->vel_is_half = 1; }

// (re)builds the packed gather copy from the planes: after a reorder, a ghost exchange or a coordinate rescale
__global__ void k_pack_q4(int n, Arrays a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a.q4[i] = make_double4(a.x[i], a.y[i], a.z[i], 0.0);
}

// barostat.update's coordinate scaling when no kick_drift follows (end of an md_step batch).
__global__ void k_scale_positions(int n, Arrays a, const Scalars *__restrict__ sc)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double mup = sc->mu_pending;
    a.x[i] = __dmul_rn(a.x[i], mup);
    a.y[i] = __dmul_rn(a.y[i], mup);
    a.z[i] = __dmul_rn(a.z[i], mup);
}

// ---- one-thread control kernels ---------------------------------------------------------------------
__global__ void k_clear_pending(Scalars *sc) { sc->mu_pending = 1.0; }

__global__ void k_after_rebuild(Scalars *sc)
{
    sc->disp_acc = 0.0;
    sc->disp_next = 0.0;
    sc->inv_scale = 1.0;
    sc->need_rebuild = 0;
    sc->out_of_box = 0;
}

__global__ void k_prepare(Scalars *sc, const Params *pr, long long n_steps, double psi)
{
    sc->steps_left = n_steps;
    sc->steps_done = 0;
    compute_controls(sc, pr, psi);
}

__global__ void k_reset_list_stats(Scalars *sc)
{
    sc->nbr_max = 0;
    sc->nbr_overflow = 0;
    sc->nbr_total = 0ull;
    sc->union_max = 0;
    sc->union_fail = 0;
}

__global__ void k_set_shift_to_vcom(Scalars *sc)
{
    sc->shift[0] = sc->vcom[0]; sc->shift[1] = sc->vcom[1]; sc->shift[2] = sc->vcom[2];
}

// ---- device-side initializer ---------------------------------------------------------
// UnitCell::{U, FCC}.initialize_particles_position (solver/src/initializer/position.rs:24-104): cell (x, y, z) has index
// x*sy*sz + y*sz + z; U puts one atom at start + (x, y, z)*l, FCC four atoms at the corner and the three face centres
// (x, y+.5, z+.5), (x+.5, y, z+.5), (x+.5, y+.5, z), in this order.  Same operations as the reference: bit-identical.
__global__ void k_init_lattice(int cells, int fcc, int sx, int sy, int sz, double x0, double y0, double z0, double l, Arrays a)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cells) return;
    const int z = c % sz, y = (c / sz) % sy, x = c / (sz * sy);
    (void)sx;
    const double fx = (double)x, fy = (double)y, fz = (double)z;
    if (!fcc) {
        a.x[c] = __dadd_rn(x0, __dmul_rn(fx, l)); a.y[c] = __dadd_rn(y0, __dmul_rn(fy, l)); a.z[c] = __dadd_rn(z0, __dmul_rn(fz, l));
        return;
    }
    const double hx = __dadd_rn(fx, 0.5), hy = __dadd_rn(fy, 0.5), hz = __dadd_rn(fz, 0.5);
    const int i = 4 * c;
    a.x[i] = __dadd_rn(x0, __dmul_rn(fx, l));     a.y[i] = __dadd_rn(y0, __dmul_rn(fy, l));     a.z[i] = __dadd_rn(z0, __dmul_rn(fz, l));
    a.x[i + 1] = __dadd_rn(x0, __dmul_rn(fx, l)); a.y[i + 1] = __dadd_rn(y0, __dmul_rn(hy, l)); a.z[i + 1] = __dadd_rn(z0, __dmul_rn(hz, l));
    a.x[i + 2] = __dadd_rn(x0, __dmul_rn(hx, l)); a.y[i + 2] = __dadd_rn(y0, __dmul_rn(fy, l)); a.z[i + 2] = __dadd_rn(z0, __dmul_rn(hz, l));
    a.x[i + 3] = __dadd_rn(x0, __dmul_rn(hx, l)); a.y[i + 3] = __dadd_rn(y0, __dmul_rn(hy, l)); a.z[i + 3] = __dadd_rn(z0, __dmul_rn(fz, l));
}

// initialize_velocities_maxwell_boltzmann (solver/src/initializer/velocity.rs:6-29): atom i < n/2 gets sigma * N(0,1) per
// component, atom i + n/2 the negated copy (an odd last atom keeps zero velocity, as in the reference's loop).  The reference
// draws from an unseeded thread_rng, so only the distribution can be matched: a counter-based generator (splitmix64 of
// (seed, atom, component)) feeds Box-Muller in f64 — reproducible for a seed and independent of the launch geometry.
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z)
{
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

__device__ __forceinline__ double standard_normal(unsigned long long seed, unsigned long long atom, int comp)
{
    const unsigned long long k = splitmix64(seed ^ splitmix64(atom * 3ull + (unsigned long long)comp));
    const unsigned long long a = splitmix64(k), b = splitmix64(k ^ 0xd1b54a32d192ed03ull);
    const double u1 = ((double)(a >> 11) + 1.0) * (1.0 / 9007199254740993.0);  // (0, 1)
    const double u2 = (double)(b >> 11) * (1.0 / 9007199254740992.0);          // [0, 1)
    return sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
}

__global__ void k_init_velocities(int n, double sigma, unsigned long long seed, Arrays a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = n / 2;
    if (i < n) {  // forces, potential and virial of a fresh State are zero
        a.fx[i] = 0.0; a.fy[i] = 0.0; a.fz[i] = 0.0; a.u[i] = 0.0; a.w[i] = 0.0;
        if (i >= 2 * half) { a.vx[i] = 0.0; a.vy[i] = 0.0; a.vz[i] = 0.0; }
    }
    if (i >= half) return;
    const double vx = sigma * standard_normal(seed, (unsigned long long)i, 0);
    const double vy = sigma * standard_normal(seed, (unsigned long long)i, 1);
    const double vz = sigma * standard_normal(seed, (unsigned long long)i, 2);
    a.vx[i] = vx; a.vy[i] = vy; a.vz[i] = vz;
    a.vx[i + half] = -vx; a.vy[i + half] = -vy; a.vz[i + half] = -vz;
}

// ---- transfer helpers -------------------------------------------------------------------------------
__global__ void k_deinterleave3(int n, const double *__restrict__ src, double *__restrict__ a,
                                double *__restrict__ b, double *__restrict__ c)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    a[i] = src[3 * (size_t)i]; b[i] = src[3 * (size_t)i + 1]; c[i] = src[3 * (size_t)i + 2];
}

__global__ void k_interleave3_unsort(int n, const double *__restrict__ a, const double *__restrict__ b,
                                     const double *__restrict__ c, const int *__restrict__ id,
                                     double *__restrict__ dst)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    size_t o = 3 * (size_t)id[p];
    dst[o] = a[p]; dst[o + 1] = b[p]; dst[o + 2] = c[p];
}

__global__ void k_unsort1(int n, const double *__restrict__ a, const int *__restrict__ id, double *__restrict__ dst)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) dst[id[p]] = a[p];
}

__global__ void k_unsort1i(int n, const int *__restrict__ a, const int *__restrict__ id, int *__restrict__ dst)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) dst[id[p]] = a[p];
}

__global__ void k_iota(int n, int *__restrict__ id)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) id[p] = p;
}


// ====================================================================================================
// Multi-GPU (1-D slabs along x in fractional coordinates).  Layout per rank:
//   [0, n_own) owned atoms in cell-sorted order | [n_own, n_own+gL) ghosts from the left neighbour | then gR from the
//   right neighbour, in the order the neighbour packed them (its own sorted order → spatially coherent).
struct Slab {
    int rank, nranks, left, right;
    double halo;  // r_list with a rounding margin
};

__device__ __forceinline__ int owner_of(double x, double Lx, int nranks) { return cell_coord(x, Lx, nranks); }

// Finalize after the all-gather of per-rank sums: every rank folds the ranks in the same order → identical
// lambda / myu / rebuild decision everywhere, deterministic for a fixed rank count.
__global__ void k_finalize_dist(const double *__restrict__ all_sums, int nranks, Scalars *sc, const Params *pr, int mode,
                                int guarded)
{
    if (guarded && halted(sc)) return;
    Sums t;
    for (int q = 0; q < NSUM; ++q) t.v[q] = 0.0;
    for (int r = 0; r < nranks; ++r) {
        for (int q = 0; q < NSUM - 1; ++q) t.v[q] += all_sums[r * NSUM + q];
        t.v[NSUM - 1] = fmax(t.v[NSUM - 1], all_sums[r * NSUM + NSUM - 1]);
    }
    finalize(sc, sc, pr, t, mode);
}

// flags for stable (scan-based) compaction: flag[i] = 1 if atom i belongs to class `want`
// class: 0 stay, 1 to the left neighbour, 2 to the right neighbour, 3 lost (moved further than one slab)
__device__ __forceinline__ int migrate_class(double x, double Lx, Slab sl)
{
    int o = owner_of(x, Lx, sl.nranks);
    if (o == sl.rank) return 0;
    if (o == sl.left) return 1;
    if (o == sl.right) return 2;
    return 3;
}

__global__ void k_flag_migrate(int n, const double *__restrict__ x, const Scalars *__restrict__ sc, Slab sl, int want,
                               int *__restrict__ flag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = migrate_class(x[i], sc->box[0], sl) == want ? 1 : 0;
}

// upload: owner selection from the full (global) arrays
__global__ void k_flag_owned(int n, const double *__restrict__ x_interleaved, double Lx, Slab sl, int *__restrict__ flag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) flag[i] = owner_of(x_interleaved[3 * (size_t)i], Lx, sl.nranks) == sl.rank ? 1 : 0;
}

// ghost candidates among the (sorted) owned atoms, selected by x CELL LAYER: side 0 = layers that reach into the halo of the
// left face, side 1 = of the right face.  A superset of the atoms within `halo` of the face (by at most one layer), and —
// because x is the slowest index of the cell sort — a prefix (side 0) / suffix (side 1) of the sorted order.
__global__ void k_flag_ghost(int n, const int *__restrict__ cell_sorted, Grid g, int layer, int side, int *__restrict__ flag)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = cell_sorted[i] / (g.nc[1] * g.nc[2]);
    flag[i] = (side == 0 ? cx <= layer : cx >= layer) ? 1 : 0;
}

// idx[pos[i]] = i for flagged i (pos = exclusive scan of flag) → ascending, deterministic
__global__ void k_compact_index(int n, const int *__restrict__ flag, const int *__restrict__ pos, int *__restrict__ idx)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flag[i]) idx[pos[i]] = i;
}

// pack planes [x | y | z | vx | vy | vz] (6*m doubles) + ids (m ints) of the atoms in idx
__global__ void k_pack_atoms(int m, const int *__restrict__ idx, Arrays a, double *__restrict__ buf, int *__restrict__ ids)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    int i = idx[k];
    buf[k] = a.x[i]; buf[m + k] = a.y[i]; buf[2 * (size_t)m + k] = a.z[i];
    buf[3 * (size_t)m + k] = a.vx[i]; buf[4 * (size_t)m + k] = a.vy[i]; buf[5 * (size_t)m + k] = a.vz[i];
    ids[k] = a.id[i];
}

__global__ void k_pack_ids(int m, const int *__restrict__ idx, const int *__restrict__ id, int *__restrict__ out)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) out[k] = id[idx[k]];
}

__global__ void k_unpack_atoms(int m, const double *__restrict__ buf, const int *__restrict__ ids, Arrays a, int at)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    int i = at + k;
    a.x[i] = buf[k]; a.y[i] = buf[m + k]; a.z[i] = buf[2 * (size_t)m + k];
    a.vx[i] = buf[3 * (size_t)m + k]; a.vy[i] = buf[4 * (size_t)m + k]; a.vz[i] = buf[5 * (size_t)m + k];
    a.fx[i] = 0.0; a.fy[i] = 0.0; a.fz[i] = 0.0; a.u[i] = 0.0; a.w[i] = 0.0;
    a.id[i] = ids[k];
}

// stayers: dst[k] = src[idx[k]] for all planes
__global__ void k_gather_atoms(int m, const int *__restrict__ idx, Arrays src, Arrays dst)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    int s = idx[k];
    dst.x[k] = src.x[s];   dst.y[k] = src.y[s];   dst.z[k] = src.z[s];
    dst.vx[k] = src.vx[s]; dst.vy[k] = src.vy[s]; dst.vz[k] = src.vz[s];
    dst.fx[k] = src.fx[s]; dst.fy[k] = src.fy[s]; dst.fz[k] = src.fz[s];
    dst.u[k] = src.u[s];   dst.w[k] = src.w[s];
    dst.id[k] = src.id[s];
}

// upload: pick the owned atoms out of the interleaved global arrays (stage = [pos | vel | force] 9n doubles optional)
__global__ void k_take_owned(int n, const int *__restrict__ flag, const int *__restrict__ pos_scan,
                             const double *__restrict__ gpos, const double *__restrict__ gvel,
                             const double *__restrict__ gforce, const double *__restrict__ gu,
                             const double *__restrict__ gw, Arrays a)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    int k = pos_scan[i];
    a.x[k] = gpos[3 * (size_t)i]; a.y[k] = gpos[3 * (size_t)i + 1]; a.z[k] = gpos[3 * (size_t)i + 2];
    a.vx[k] = gvel[3 * (size_t)i]; a.vy[k] = gvel[3 * (size_t)i + 1]; a.vz[k] = gvel[3 * (size_t)i + 2];
    a.fx[k] = gforce ? gforce[3 * (size_t)i] : 0.0;
    a.fy[k] = gforce ? gforce[3 * (size_t)i + 1] : 0.0;
    a.fz[k] = gforce ? gforce[3 * (size_t)i + 2] : 0.0;
    a.u[k] = gu ? gu[i] : 0.0;
    a.w[k] = gw ? gw[i] : 0.0;
    a.id[k] = i;
}

// per-step halo: positions of the atoms in idx → buf [x | y | z] (3*m doubles); and the inverse on the receiver
__global__ void k_pack_halo(int m, const int *__restrict__ idx, const double *__restrict__ x, const double *__restrict__ y,
                            const double *__restrict__ z, double *__restrict__ buf, const Scalars *__restrict__ sc,
                            int guarded)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    if (guarded && halted(sc)) return;
    int i = idx[k];
    buf[k] = x[i]; buf[m + k] = y[i]; buf[2 * (size_t)m + k] = z[i];
}

__global__ void k_unpack_halo(int m, const double *__restrict__ buf, double *__restrict__ x, double *__restrict__ y,
                              double *__restrict__ z, int at, const Scalars *__restrict__ sc, int guarded,
                              double4 *__restrict__ q4)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    if (guarded && halted(sc)) return;
    const double px = buf[k], py = buf[m + k], pz = buf[2 * (size_t)m + k];
    x[at + k] = px; y[at + k] = py; z[at + k] = pz;
    if (q4) q4[at + k] = make_double4(px, py, pz, 0.0);
}

// local cell index: x is measured from (slab lower face - halo), unwrapped periodically; y, z as in the global grid
__device__ __forceinline__ int cell_coord_local_x(double x, double Lx, Slab sl, double extent, int nc)
{
    const double xo = Lx * ((double)sl.rank / (double)sl.nranks) - sl.halo;
    double d = x - xo;
    if (d < 0.0) d += Lx;
    else if (d >= Lx) d -= Lx;
    int c = (int)(d / extent * (double)nc);
    return min(max(c, 0), nc - 1);
}

__device__ __forceinline__ double slab_extent(double Lx, Slab sl)
{
    return Lx / (double)sl.nranks + 2.0 * sl.halo;
}

__global__ void k_cell_count_dist(int n, int first, Arrays a, const Scalars *__restrict__ sc, Grid g, Slab sl,
                                  int *__restrict__ cell_of, int *__restrict__ cell_cnt)
{
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int i = first + k;
    const double Lx = sc->box[0];
    int cx = cell_coord_local_x(a.x[i], Lx, sl, slab_extent(Lx, sl), g.nc[0]);
    int cy = cell_coord(a.y[i], sc->box[1], g.nc[1]);
    int cz = cell_coord(a.z[i], sc->box[2], g.nc[2]);
    int c = (cx * g.nc[1] + cy) * g.nc[2] + cz;
    cell_of[k] = c;
    atomicAdd(&cell_cnt[c], 1);
}

// k_sort_cells for the ghost table: order[] holds ghost receive indices, keys are their global ids
__global__ void k_sort_cells_offset(int ncell, const int *__restrict__ cell_start, const int *__restrict__ id, int first,
                                    int *__restrict__ order)
{
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    int s = cell_start[c], e = cell_start[c + 1];
    for (int a = s + 1; a < e; ++a) {
        int item = order[a];
        int key = id[first + item];
        int b = a - 1;
        while (b >= s && id[first + order[b]] > key) {
            order[b + 1] = order[b];
            --b;
        }
        order[b + 1] = item;
    }
}

// K2 for a slab: stencil is periodic in y, z and open in x (the halo supplies the partners beyond the faces);
// partners come from the owned cell table and, through ghost_order, from the ghost cell table.
template <bool SORT_BY_ID>
__global__ void __launch_bounds__(128) k_build_list_dist(int n_own, Arrays a, const int *__restrict__ cell_sorted,
                                                         const int *__restrict__ cell_start,
                                                         const int *__restrict__ ghost_start,
                                                         const int *__restrict__ ghost_order, Scalars *sc, Grid g,
                                                         double r_list, double r2_list, int *__restrict__ nbr,
                                                         int *__restrict__ nbr_cnt)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    if (p < n_own) {
        const double Lx = sc->box[0], Ly = sc->box[1], Lz = sc->box[2];
        const double hx = Lx / 2.0, hy = Ly / 2.0, hz = Lz / 2.0;
        const double xi = a.x[p], yi = a.y[p], zi = a.z[p];
        const int ncx = g.nc[0], ncy = g.nc[1], ncz = g.nc[2];
        int c = cell_sorted[p];
        int cz = c % ncz;
        int cy = (c / ncz) % ncy;
        int cx = c / (ncz * ncy);
        int w = 2 * g.nsub + 1;
        int loy, ny;
        if (ncy >= w) { loy = cy - g.nsub; ny = w; } else { loy = 0; ny = ncy; }
        int z0a, z1a, z0b = 0, z1b = 0;
        if (ncz >= w) {
            int lo = cz - g.nsub, hi = cz + g.nsub + 1;
            if (lo < 0) { z0a = 0; z1a = hi; z0b = lo + ncz; z1b = ncz; }
            else if (hi > ncz) { z0a = lo; z1a = ncz; z0b = 0; z1b = hi - ncz; }
            else { z0a = lo; z1a = hi; }
        } else { z0a = 0; z1a = ncz; }
        for (int qx = max(cx - g.nsub, 0); qx <= min(cx + g.nsub, ncx - 1); ++qx) {
            for (int ib = 0; ib < ny; ++ib) {
                int qy = loy + ib;
                qy += (qy < 0) ? ncy : 0;
                qy -= (qy >= ncy) ? ncy : 0;
                const int base = (qx * ncy + qy) * ncz;
#pragma unroll 1
                for (int pass = 0; pass < 4; ++pass) {  // owned run a, owned run b, ghost run a, ghost run b
                    const bool ghost = pass >= 2;
                    const int z0 = (pass & 1) ? z0b : z0a, z1 = (pass & 1) ? z1b : z1a;
                    if (z1 <= z0) continue;
                    const int *__restrict__ tab = ghost ? ghost_start : cell_start;
                    const int s = tab[base + z0], e = tab[base + z1];
                    for (int t = s; t < e; ++t) {
                        const int q = ghost ? n_own + ghost_order[t] : t;
                        double rx = min_image(__dsub_rn(a.x[q], xi), Lx, hx);
                        if (fabs(rx) > r_list) continue;
                        double ry = min_image(__dsub_rn(a.y[q], yi), Ly, hy);
                        double rz = min_image(__dsub_rn(a.z[q], zi), Lz, hz);
                        double r2 = __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz));
                        if (r2 > r2_list || q == p) continue;
                        if (cnt < g.cap) nbr[(size_t)cnt * g.npad + p] = q;
                        ++cnt;
                    }
                }
            }
        }
        nbr_cnt[p] = min(cnt, g.cap);
        if (SORT_BY_ID && cnt <= g.cap) {
            for (int s1 = 1; s1 < cnt; ++s1) {
                int item = nbr[(size_t)s1 * g.npad + p];
                int key = a.id[item];
                int b = s1 - 1;
                while (b >= 0 && a.id[nbr[(size_t)b * g.npad + p]] > key) {
                    nbr[(size_t)(b + 1) * g.npad + p] = nbr[(size_t)b * g.npad + p];
                    --b;
                }
                nbr[(size_t)(b + 1) * g.npad + p] = item;
            }
        }
    }
    int wmax = cnt;
    unsigned int wsum = (unsigned int)cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    }
    if ((threadIdx.x & 31) == 0 && wsum) {
        atomicMax(&sc->nbr_max, wmax);
        atomicAdd(&sc->nbr_total, (unsigned long long)wsum);
        if (wmax > g.cap) atomicExch(&sc->nbr_overflow, 1);
    }
}

__global__ void k_fill_int(int n, int *p, int v)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// download helper: owned atoms' planes → interleaved staging in the CURRENT (sorted) order + ids
__global__ void k_interleave3(int n, const double *__restrict__ a, const double *__restrict__ b,
                              const double *__restrict__ c, double *__restrict__ dst)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    dst[3 * (size_t)p] = a[p]; dst[3 * (size_t)p + 1] = b[p]; dst[3 * (size_t)p + 2] = c[p];
}

}  // namespace md

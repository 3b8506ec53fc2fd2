// md_integrate.cuh — K4: k_kick_drift, the fused one-kernel step, state upload/download helpers, initializer.
// Part of md_kernels.cuh (included from there, in order; one translation unit).
#pragma once

namespace md {

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

}  // namespace md

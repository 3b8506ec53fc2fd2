// md_force.cuh — K3: pair-force kernels (k_force and its loops, k_force_sparse, k_force_coop).
// Part of md_kernels.cuh (included from there, in order; one translation unit).
#pragma once

namespace md {

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
        const double inv = rcp_nr(r2);  // <= 1.4e-14 relative (the FAST bar is 1e-10); the IEEE division's slow path is ~4x the code
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
template <int B>
struct SumsSmemT {
    double v[NSUM][B];
};
using SumsSmem = SumsSmemT<FORCE_BLOCK>;

// nh (uniform): also accumulate the COM/thermal sums of u', which only Nose-Hoover's second psi update reads.
template <typename SS>
__device__ __forceinline__ void finish_atom(SS &ss, const PairAcc &f, double &vx, double &vy, double &vz,
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

// Neighbour loop of one atom pair, dilute systems (FAST mode): branchy pair term, plane gathers; the next row of partner
// indices is prefetched while the current one is in flight.
__device__ __forceinline__ void neighbour_loop(PairAcc &f0, PairAcc &f1, const Arrays &a, const int2 *__restrict__ row,
                                               size_t stride, int last_row, int2 C, int i0, double2 X, double2 Y,
                                               double2 Z, const LjConst &c, const ForceConsts &fc, int2 Ja)
{
    // Ja = row[0]: it exists for every atom (cap >= 8) and the caller fetched it together with the atom's own data
    const double *__restrict__ px = a.x, *__restrict__ py = a.y, *__restrict__ pz = a.z;
    const int kmax = max(C.x, C.y);
    for (int k = 0; k < kmax; ++k) {
        const int2 Na = row[min(k + 1, last_row) * stride];
        const bool a0 = k < C.x, a1 = k < C.y;
        const int ja0 = a0 ? Ja.x : i0, ja1 = a1 ? Ja.y : i0;
        const double xa0 = px[ja0], ya0 = py[ja0], za0 = pz[ja0];
        const double xa1 = px[ja1], ya1 = py[ja1], za1 = pz[ja1];
        pair_fast_branchy(f0, a0, xa0, ya0, za0, X.x, Y.x, Z.x, c, fc);
        pair_fast_branchy(f1, a1, xa1, ya1, za1, X.y, Y.y, Z.y, c, fc);
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

// shared by the force kernels: guarded early-out, phase clock, wait for the neighbours' ghosts (peer-memory path)
__device__ __forceinline__ bool force_prologue(int do_step, Scalars *sc, const Peers *peers)
{
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

// Two consecutive atoms per thread: every plane access is one 128-bit transaction, all of a pair's loads are issued
// before the first use, and the neighbour loop advances both lists together (independent gather chains) with the
// next rows of partner indices prefetched while the current ones are in flight.
//   EXACT:  the reference's arithmetic and summation order (bit-identical), any density
//   MASKED: FAST mode, dense systems (branch-free pair term, packed gathers, index ring in shared memory)
//   else:   FAST mode, dilute systems (branchy pair term behind the distance test)
template <bool EXACT, bool MASKED>
__global__ void __launch_bounds__(FORCE_BLOCK, (EXACT || MASKED) ? MD_FORCE_MINB : MD_FORCE_MINB_DILUTE)
    k_force(int n, Arrays a, const int *__restrict__ nbr, const int *__restrict__ nbr_cnt, int npad, int cap,
            double *__restrict__ partials, Scalars *sc, const Params *__restrict__ pr, int do_step,
            const ForceConsts fc, const Peers *peers)
{
    // do_step bits: 1 = MD step (both half-kicks fused in), 2 = multi-GPU (publish rank sums only), 4 = guarded,
    //               8 = multi-GPU over peer memory (sums exchanged and finalized here), 16 = wait for the neighbours' ghosts
    if (!force_prologue(do_step, sc, peers)) return;
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
            C = reinterpret_cast<const int2 *>(nbr_cnt)[t];
            J0 = row[0];
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
#define MD_DENSE_CALL(W, U) neighbour_loop_dense<W, U>(f0, f1, a.q4, row, stride, C, i0, X, Y, Z, c, fc, ring, n)
                if (wrap) {
                    if (uw == 0) MD_DENSE_CALL(true, 0); else if (uw == 1) MD_DENSE_CALL(true, 1); else MD_DENSE_CALL(true, 2);
                } else {
                    if (uw == 0) MD_DENSE_CALL(false, 0); else if (uw == 1) MD_DENSE_CALL(false, 1); else MD_DENSE_CALL(false, 2);
                }
#undef MD_DENSE_CALL
            } else {
                neighbour_loop(f0, f1, a, row, stride, last_row, C, i0, X, Y, Z, c, fc, J0);
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
    Sums s;
#pragma unroll
    for (int q = 0; q < NSUM; ++q) s.v[q] = ss.v[q][threadIdx.x];
    block_reduce<FORCE_BLOCK>(s);
    grid_reduce_finalize<FORCE_BLOCK>(s, partials, sc, pr,
                                      (do_step & 1 ? FIN_STEP : 0) | (do_step & 2 ? FIN_DIST : 0) | (do_step & 8 ? FIN_P2P : 0),
                                      peers);
}

}  // namespace md

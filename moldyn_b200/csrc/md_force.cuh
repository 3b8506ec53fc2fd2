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
// ncu on k_force<.., MASKED> (profiles/r01_ncu_c5_v10): 288 k global data-pipe wavefronts per SM in 360 k active cycles
// (l1tex__data_pipe_lsu_wavefronts 74 % of peak on average, 89 % on the busiest SM) — the per-thread Verlet loop is bound by
// the L1 data pipe, not by FP64 (32 % busy): lane l gathers partner k of ITS atom, the lanes hit ~28 different 128-byte
// lines per instruction, one wavefront each.  Here the 32 lanes of a warp work on ONE atom at a time: lane l takes list
// entries l, l + 32, ...  The list is stored atom-major (k_transpose_list), so the index read is one coalesced line, and
// since a list is the concatenation of ascending index runs (one per stencil cell run) consecutive entries are mostly
// consecutive atoms: up to four partners share a 128-byte line of the packed copy q4.  Measured: ~17 wavefronts per gather
// instead of ~28 — the distance filter fragments the runs — bought with 16 % more instructions; slower overall.  The per-lane partial forces are folded with xor-shuffles (fixed order: deterministic) and handed to the lane
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

}  // namespace md

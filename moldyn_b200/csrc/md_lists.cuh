// md_lists.cuh — K2: Verlet-skin neighbour lists (per atom, and union lists per atom pair).
// Part of md_kernels.cuh (included from there, in order; one translation unit).
#pragma once

namespace md {

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

}  // namespace md

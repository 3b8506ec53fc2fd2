// md_lists.cuh — K2: Verlet-skin neighbour lists (per atom).
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
        int cx, cy, cz;
        cell_decode(g, cell_sorted[p], cx, cy, cz);
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
                const int base = col_index(g, qx, qy) * ncz;
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

}  // namespace md

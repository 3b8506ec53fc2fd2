// md_cells.cuh — K1: cell index, histogram, scan, counting-sort scatter, per-cell sort, reorder.
// Part of md_kernels.cuh (included from there, in order; one translation unit).
#pragma once

namespace md {

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
    int c = cell_index(g, cx, cy, cz);
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

}  // namespace md

// md_dist_kernels.cuh — multi-GPU slab kernels (ownership, migration, halo packing, distributed list build).
// Part of md_kernels.cuh (included from there, in order; one translation unit).
#pragma once

namespace md {

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
                                                         int *__restrict__ nbr_cnt, int *__restrict__ nbr_ghost)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    int cnt = 0;
    if (p < n_own) {
        int any_ghost = 0;  // does the list hold a ghost?  (the persistent step loop computes such atoms last)
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
                        any_ghost |= ghost ? 1 : 0;
                    }
                }
            }
        }
        nbr_cnt[p] = min(cnt, g.cap);
        nbr_ghost[p] = any_ghost;
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

// md_tile.cuh — K2 + K3 for dense systems (FAST mode, one GPU): brick tiles staged in shared memory, brick-local 16-bit
// neighbour lists, warp-cooperative pair loop with shuffle accumulation.
// Part of md_kernels.cuh (included from there, in order; one translation unit).
#pragma once

namespace md {

// Why (profiles/r01_ncu_c5_v10_k_force.txt): the per-thread Verlet loop of k_force<.., MASKED> gathers every partner from
// global memory — one L1 wavefront per lane and partner, l1tex data pipe 74 % busy on average and 89 % on the busiest SM,
// FP64 pipe a third busy.  Here a thread block owns a BRICK of the cell grid (4 x 4 columns x bz cells, a few hundred atoms)
// and stages the brick plus the two-cell shell around it — every possible partner of its atoms, ~8 x the brick — in shared
// memory once: the cell sort numbers the (x, y) columns brick by brick (Grid::brick), so the shell is 64 columns x (one or
// two) contiguous z-runs of the sorted planes.  The lists hold 16-bit indices into the shell, stored atom-major; eight
// lanes of a warp walk ONE atom's list together (shared-memory gathers of mostly consecutive slots: conflict-free), four
// independent pair terms per lane in flight, and fold their partial forces with xor-shuffles in a fixed order
// (deterministic).  Periodic images are resolved when the shell is staged, so the pair loop has no minimum-image step.
// The K2 kernel stages the same shell (same code, same local indices) and evaluates the reference's predicate
// (potential.rs:181-204 widened by the skin) in the reference's own operation order on the unshifted coordinates.
constexpr int TILE_BLOCK = 256;
constexpr int TILE_WARPS = TILE_BLOCK / 32;
constexpr int TILE_WIN = 8;                          // window columns per dimension: 4 own + 2 on either side
constexpr int TILE_COLS = TILE_WIN * TILE_WIN;       // 64
constexpr double TILE_FAR = 1.0e7;                   // coordinate of the padding slot [nm]
constexpr int TILE_RUNS = 2 * TILE_COLS;             // per column: the unwrapped z-run and the periodically wrapped one
constexpr int TILE_MIN_CELLS = 8;                    // cells per dimension the brick order needs (a window never meets itself)

struct TileTab {
    int run_src[TILE_RUNS];   // sorted index of the run's first atom
    int run_len[TILE_RUNS];
    int run_loc[TILE_RUNS];   // shell slot of the run's first atom
    int col_base[TILE_COLS];  // first cell of the window column in the cell table
    double shx[TILE_COLS], shy[TILE_COLS];  // image shift of the window column (-L, 0, +L)
    double shz;               // image shift of the wrapped z-runs (odd run index)
    double ctr[3];            // centre of the brick's nominal extent: staged atoms are re-imaged next to it
    int own_src[16], own_loc[16], own_pref[17];
    int n_shell, n_own;
    int zlo, zhi;             // own z cells [zlo, zhi)
};

// Brick `id` → its tables.  Block-wide (>= 64 threads); ends with a barrier.  `measure`: sizes only.
__device__ __forceinline__ void tile_setup(TileTab &T, const Grid &g, const int *__restrict__ cell_start, const Scalars *sc,
                                           int id)
{
    const int t = threadIdx.x;
    const int ncx = g.nc[0], ncy = g.nc[1], ncz = g.nc[2];
    const int bzc = id % g.nbz, bxy = id / g.nbz, by = bxy % g.nby, bx = bxy / g.nby;
    const int zlo = g.bz * bzc, zhi = min(g.bz * (bzc + 1), ncz);
    const int uz0 = zlo - 2, uz1 = zhi + 2;  // window along z, unwrapped cell coordinates
    if (t < TILE_COLS) {
        const int wx = t >> 3, wy = t & 7;
        const int ux = 4 * bx - 2 + wx, uy = 4 * by - 2 + wy;
        const int sxc = ux < 0 ? -1 : (ux >= ncx ? 1 : 0), syc = uy < 0 ? -1 : (uy >= ncy ? 1 : 0);
        const int cx = ux - sxc * ncx, cy = uy - syc * ncy;
        const int base = col_index(g, cx, cy) * ncz;
        T.col_base[t] = base;
        T.shx[t] = sxc < 0 ? -sc->box[0] : (sxc > 0 ? sc->box[0] : 0.0);
        T.shy[t] = syc < 0 ? -sc->box[1] : (syc > 0 ? sc->box[1] : 0.0);
        const int zA0 = max(uz0, 0), zA1 = min(uz1, ncz);
        int zB0 = 0, zB1 = 0;
        if (uz0 < 0) { zB0 = uz0 + ncz; zB1 = ncz; }
        else if (uz1 > ncz) { zB0 = 0; zB1 = uz1 - ncz; }
        const int sA = cell_start[base + zA0], eA = cell_start[base + zA1];
        const int sB = zB1 > zB0 ? cell_start[base + zB0] : 0, eB = zB1 > zB0 ? cell_start[base + zB1] : 0;
        T.run_src[2 * t] = sA; T.run_len[2 * t] = eA - sA;
        T.run_src[2 * t + 1] = sB; T.run_len[2 * t + 1] = eB - sB;
        if (t == 0) {
            T.shz = uz0 < 0 ? -sc->box[2] : sc->box[2];
            T.zlo = zlo; T.zhi = zhi;
            T.ctr[0] = (4 * bx + 2) * (sc->box[0] / ncx);
            T.ctr[1] = (4 * by + 2) * (sc->box[1] / ncy);
            T.ctr[2] = 0.5 * (zlo + zhi) * (sc->box[2] / ncz);
        }
    }
    __syncthreads();
    if (t < 32) {  // exclusive scan of the 128 run lengths: four runs per lane
        int c[4], sum = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) { c[k] = T.run_len[4 * t + k]; sum += c[k]; }
        int inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (t >= o) inc += v;
        }
        int off = inc - sum;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            T.run_loc[4 * t + k] = off;
            off += c[k];
        }
        if (t == 31) T.n_shell = inc;
    }
    __syncthreads();
    if (t < 16) {
        const int lx = t >> 2, ly = t & 3;
        const bool exists = 4 * bx + lx < ncx && 4 * by + ly < ncy;
        const int wcol = (lx + 2) * TILE_WIN + (ly + 2), rA = 2 * wcol;
        const int base = T.col_base[wcol];
        const int s = exists ? cell_start[base + zlo] : 0, e = exists ? cell_start[base + zhi] : 0;
        T.own_src[t] = s;
        T.own_loc[t] = T.run_loc[rA] + (s - T.run_src[rA]);
        T.own_pref[t + 1] = e - s;  // counts for now
    }
    __syncthreads();
    if (t == 0) {
        int acc = 0;
        T.own_pref[0] = 0;
        for (int k = 0; k < 16; ++k) { acc += T.own_pref[k + 1]; T.own_pref[k + 1] = acc; }
        T.n_own = acc;
    }
    __syncthreads();
}

// Largest shell and brick of the current sort: sizes the tile kernels' shared memory.
__global__ void __launch_bounds__(64) k_tile_measure(Grid g, const int *__restrict__ cell_start, Scalars *sc)
{
    __shared__ TileTab T;
    tile_setup(T, g, cell_start, sc, blockIdx.x);
    if (threadIdx.x == 0) {
        atomicMax(&sc->tile_shell_max, T.n_shell);
        atomicMax(&sc->tile_own_max, T.n_own);
    }
}

__global__ void k_tile_reset(Scalars *sc)
{
    sc->tile_shell_max = 0;
    sc->tile_own_max = 0;
}

// Stages the shell: the x, y, z planes of every run → sp[3 * slot + {0, 1, 2}] (array of structures: a partner costs the pair
// loop ONE address computation and three LDS.64 at immediate offsets; consecutive slots are conflict-free — 24-byte stride).
//   image = false (list builder): the stored coordinates as they are — the builder applies the reference's own (x_q - x_i) -+ L.
//   image = true (force kernel): every atom is placed next to the brick — the run's periodic shift, then one more box length
//     if the atom has crossed a box face since the lists were built (the drift wraps coordinates into [0, L); relative to the
//     brick centre the true image is the one within half a box).  The pair loop then needs no minimum-image step at all.
// The block's threads copy the runs themselves, coalesced within a run.  (A cp.async.bulk — TMA — copy per run and plane into
// separate planes was measured first: k_force_tile 263.3 vs 261.0 us on C5, no gain — a shell is ~200 runs of ~200 bytes, too
// small and too irregular for bulk copies to pay — and the planes cost the pair loop two more address computations per
// partner.  profiles/r02_bench_c5_tile_v1.txt)
__device__ __forceinline__ double tile_image(double x, double shift, double centre, double L)
{
    x += shift;
    const double d = x - centre, h = 0.5 * L;
    return d > h ? x - L : (d < -h ? x + L : x);
}

__device__ __forceinline__ void tile_stage(const TileTab &T, const Arrays &a, const Scalars *sc, double *sp, bool image)
{
    // One shell slot per thread and trip, four trips in flight: the slot's run is found by bisection over the runs' first
    // slots (shared memory), so all of a thread's global loads are independent of each other.  (Walking the runs one after
    // the other cost a global-load latency per run, 16 per warp: ~15 % of the first versions' kernel time sat in the
    // prologue, profiles/r02_ncu_c5_force_tile_v3.txt.)
    const double Lx = sc->box[0], Ly = sc->box[1], Lz = sc->box[2];
    const int n_shell = T.n_shell;
    for (int s0 = threadIdx.x; s0 < n_shell; s0 += 4 * TILE_BLOCK) {
        int r[4], g[4];
        double x[4], y[4], z[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int s = s0 + u * TILE_BLOCK;
            int lo = 0;  // last run whose first slot is <= s (empty runs share their successor's first slot: skipped by >=)
#pragma unroll
            for (int step = TILE_RUNS / 2; step > 0; step >>= 1)
                if (lo + step < TILE_RUNS && T.run_loc[lo + step] <= s) lo += step;
            r[u] = lo;
            g[u] = s < n_shell ? T.run_src[lo] + (s - T.run_loc[lo]) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { x[u] = a.x[g[u]]; y[u] = a.y[g[u]]; z[u] = a.z[g[u]]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int s = s0 + u * TILE_BLOCK;
            if (s < n_shell) {
                const double dx = T.shx[r[u] >> 1], dy = T.shy[r[u] >> 1], dz = (r[u] & 1) ? T.shz : 0.0;
                double *o = sp + 3 * s;
                o[0] = image ? tile_image(x[u], dx, T.ctr[0], Lx) : x[u];
                o[1] = image ? tile_image(y[u], dy, T.ctr[1], Ly) : y[u];
                o[2] = image ? tile_image(z[u], dz, T.ctr[2], Lz) : z[u];
            }
        }
    }
    // Slot n_shell is nobody: far outside every cutoff.  The builder pads each list to a multiple of 32 entries with it, so the
    // pair loop needs no per-entry "is this entry real" test — a padding entry fails the range test like any distant partner.
    if (threadIdx.x == 0) { sp[3 * n_shell] = TILE_FAR; sp[3 * n_shell + 1] = TILE_FAR; sp[3 * n_shell + 2] = TILE_FAR; }
    __syncthreads();
}

// The builder's view of the shell: single-precision coordinates relative to the brick centre, the run's periodic shift folded
// in, one 16-byte record per slot (ONE LDS.128 per candidate).  They only PRE-FILTER: a candidate whose single-precision
// distance² is not within TILE_BAND of the list radius² is decided there (coordinates within ~4 nm of the centre: each is
// off by < 2.4e-7 nm, distance² by < 3e-6 nm² — TILE_BAND is 30 times that); the few inside the band are decided by the
// reference's own double-precision operations on the stored coordinates.
constexpr float TILE_BAND = 1.0e-4f;
__device__ __forceinline__ void tile_stage_f32(const TileTab &T, const Arrays &a, float4 *sf)
{
    const int n_shell = T.n_shell;
    for (int s0 = threadIdx.x; s0 < n_shell; s0 += 4 * TILE_BLOCK) {
        int r[4], g[4];
        double x[4], y[4], z[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int s = s0 + u * TILE_BLOCK;
            int lo = 0;
#pragma unroll
            for (int step = TILE_RUNS / 2; step > 0; step >>= 1)
                if (lo + step < TILE_RUNS && T.run_loc[lo + step] <= s) lo += step;
            r[u] = lo;
            g[u] = s < n_shell ? T.run_src[lo] + (s - T.run_loc[lo]) : 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { x[u] = a.x[g[u]]; y[u] = a.y[g[u]]; z[u] = a.z[g[u]]; }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int s = s0 + u * TILE_BLOCK;
            if (s < n_shell) {
                const double dx = T.shx[r[u] >> 1], dy = T.shy[r[u] >> 1], dz = (r[u] & 1) ? T.shz : 0.0;
                sf[s] = make_float4((float)((x[u] + dx) - T.ctr[0]), (float)((y[u] + dy) - T.ctr[1]),
                                    (float)((z[u] + dz) - T.ctr[2]), 0.f);
            }
        }
    }
    __syncthreads();
}

// brick atom a (0 <= a < n_own) → own column, shell slot, sorted index
__device__ __forceinline__ void tile_locate(const TileTab &T, int a, int &oc, int &slot, int &gi)
{
    oc = 0;
#pragma unroll
    for (int step = 8; step > 0; step >>= 1)
        if (a >= T.own_pref[oc + step]) oc += step;
    const int k = a - T.own_pref[oc];
    slot = T.own_loc[oc] + k;
    gi = T.own_src[oc] + k;
}

// K2, tile form.  One block per brick, ONE THREAD PER ATOM: the thread walks its atom's 25 stencil columns — per column one
// contiguous range of shell slots for the unwrapped z cells and one for the periodically wrapped ones — out of shared memory,
// evaluating the reference's predicate (potential.rs:181-204 widened by the skin) in the reference's own operation order.
// Hits collect in a 32-entry buffer of the thread (64 bytes of shared memory) that is written to the atom's row as two
// sectors whenever it fills: the list is ascending in (column, slot), 16-bit shell slots, atom-major (nbrT[i * cap + ..]),
// each 32-entry chunk stored in the order the pair loop's lanes read it (tile_chunk_pos) and the last one padded.
// (Two warp-cooperative forms came first — lanes over a column's candidates, then one lane per column with bit masks: 1.36
// and 0.95 ms on C5, 2800 and 2260 instructions per atom, half the lanes idle; profiles/r02_ncu_c5_build_tile_v2.txt.)
constexpr int TILE_BUF = 32;
// Position of hit h (0..31) of a 32-entry chunk in memory: lane l8 of the pair loop reads positions 4*l8 .. 4*l8+3 as one 8-byte
// word and wants hits l8, l8+8, l8+16, l8+24 there — at a fixed u the eight lanes of an atom then touch CONSECUTIVE hits, i.e.
// neighbouring shell slots (few bank conflicts; with hits 4*l8+u instead the kernel lost what the shorter loop gained,
// profiles/r02_ncu_c5_force_tile_v5.txt: short_scoreboard 1.1 -> 1.95).
__host__ __device__ constexpr int tile_chunk_pos(int h) { return ((h & 7) << 2) | (h >> 3); }

__global__ void __launch_bounds__(TILE_BLOCK) k_build_tile(Grid g, Arrays a, const int *__restrict__ cell_start,
                                                           const int *__restrict__ cell_sorted, Scalars *sc,
                                                           double r2_list, unsigned short *__restrict__ nbrT, int cap,
                                                           int *__restrict__ nbr_cnt, const int *__restrict__ brick_order,
                                                           int sh_cap)
{
    extern __shared__ __align__(16) unsigned char tile_smem[];
    float4 *sf = reinterpret_cast<float4 *>(tile_smem);
    uint4 *bufs = reinterpret_cast<uint4 *>(sf + (size_t)sh_cap);  // TILE_BLOCK buffers of TILE_BUF 16-bit entries
    __shared__ TileTab T;
    tile_setup(T, g, cell_start, sc, brick_order[blockIdx.x]);
    tile_stage_f32(T, a, sf);
    const int ncz = g.nc[2];
    unsigned short *mybuf = reinterpret_cast<unsigned short *>(bufs + 4 * threadIdx.x);
    const float r2_lo = (float)r2_list - TILE_BAND, r2_hi = (float)r2_list + TILE_BAND;
    int wmax = 0;
    unsigned long long wsum = 0ull;
    for (int ai = threadIdx.x; ai < T.n_own; ai += TILE_BLOCK) {
        int oc, slot, gi;
        tile_locate(T, ai, oc, slot, gi);
        const double xi = a.x[gi], yi = a.y[gi], zi = a.z[gi];
        const float4 fi = sf[slot];
        const int cz = cell_sorted[gi] % ncz;
        const int lx = oc >> 2, ly = oc & 3;
        const int za = max(cz - 2, 0), zb = min(cz + 3, ncz);
        int zc = 0, zd = 0;  // wrapped cells of the stencil
        if (cz - 2 < 0) { zc = cz - 2 + ncz; zd = ncz; }
        else if (cz + 3 > ncz) { zc = 0; zd = cz + 3 - ncz; }
        uint4 *__restrict__ out = reinterpret_cast<uint4 *>(nbrT + (size_t)gi * cap);
        int cnt = 0;
        auto d2_of = [&](int q) {
            const float4 f = sf[q];
            const float rx = f.x - fi.x, ry = f.y - fi.y, rz = f.z - fi.z;
            return fmaf(rz, rz, fmaf(ry, ry, rx * rx));
        };
        auto take = [&](int q) {
            mybuf[tile_chunk_pos(cnt & (TILE_BUF - 1))] = (unsigned short)q;
            ++cnt;
            if ((cnt & (TILE_BUF - 1)) == 0 && cnt <= cap) {
                const int chunk = (cnt >> 5) - 1;
                asm volatile("" ::: "memory");  // (the buffer is written as 16-bit words and read back as four 16-byte ones)
#pragma unroll
                for (int v = 0; v < 4; ++v) out[4 * chunk + v] = bufs[4 * threadIdx.x + v];
            }
        };
        // slot q of a run whose slot s is atom gs of the sorted order: the reference's operations in the reference's order on
        // the stored coordinates — (x_q - x_i) -+ L (adding 0.0 is exact), then norm².  r2_list is the largest double whose
        // square root is <= r_list: r2 <= r2_list IS the reference's norm(r) <= r_list.
        auto exact_hit = [&](int q, int s, int gs, double dx, double dy, double dz) {
            const int gq = gs + (q - s);
            const double rx = __dadd_rn(__dsub_rn(a.x[gq], xi), dx);
            const double ry = __dadd_rn(__dsub_rn(a.y[gq], yi), dy);
            const double rz = __dadd_rn(__dsub_rn(a.z[gq], zi), dz);
            return __dadd_rn(__dadd_rn(__dmul_rn(rx, rx), __dmul_rn(ry, ry)), __dmul_rn(rz, rz)) <= r2_list;
        };
        auto scan = [&](int s, int e, int gs, double dx, double dy, double dz) {
            // four candidates per trip: their distance chains are independent (a thread's scan is otherwise one long
            // dependency chain — 'wait' was the first version's dominant stall); hits are taken in slot order
            auto decide = [&](int q, float d2) {
                if (d2 < r2_hi && q != slot && (d2 <= r2_lo || exact_hit(q, s, gs, dx, dy, dz))) take(q);
            };
            int q = s;
            for (; q + 4 <= e; q += 4) {
                const float d0 = d2_of(q), d1 = d2_of(q + 1), d2 = d2_of(q + 2), d3 = d2_of(q + 3);
                decide(q, d0); decide(q + 1, d1); decide(q + 2, d2); decide(q + 3, d3);
            }
            for (; q < e; ++q) decide(q, d2_of(q));
        };
        for (int col = 0; col < 25; ++col) {
            const int wcol = (lx + col / 5) * TILE_WIN + (ly + col % 5);
            const int base = T.col_base[wcol];
            const int ca = cell_start[base + za], cb = cell_start[base + zb];
            const int sA = T.run_loc[2 * wcol] + (ca - T.run_src[2 * wcol]);
            const double dx = T.shx[wcol], dy = T.shy[wcol];
            scan(sA, sA + (cb - ca), ca, dx, dy, 0.0);
            if (zd > zc) {
                const int cc = cell_start[base + zc], cd = cell_start[base + zd];
                const int sB = T.run_loc[2 * wcol + 1] + (cc - T.run_src[2 * wcol + 1]);
                scan(sB, sB + (cd - cc), cc, dx, dy, T.shz);
            }
        }
        // pad to a multiple of 32 entries with the nobody slot (cap is a multiple of 32: the row has room), which also
        // flushes the last, partly filled buffer
        if (cnt <= cap && (cnt & (TILE_BUF - 1)) != 0) {
            for (int h = cnt & (TILE_BUF - 1); h < TILE_BUF; ++h) mybuf[tile_chunk_pos(h)] = (unsigned short)T.n_shell;
            const int chunk = cnt >> 5;
            asm volatile("" ::: "memory");
#pragma unroll
            for (int v = 0; v < 4; ++v) out[4 * chunk + v] = bufs[4 * threadIdx.x + v];
        }
        nbr_cnt[gi] = min(cnt, cap);
        wmax = max(wmax, cnt);
        wsum += (unsigned long long)cnt;
    }
    // statistics: max / total / overflow (integer atomics — order-independent results)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
        wsum += __shfl_xor_sync(0xffffffffu, wsum, o);
    }
    if ((threadIdx.x & 31) == 0 && wsum) {
        atomicMax(&sc->nbr_max, wmax);
        atomicAdd(&sc->nbr_total, wsum);
        if (wmax > cap) atomicExch(&sc->nbr_overflow, 1);
    }
}

// Introspection (md_neighbour_lists): the brick-local lists as sorted indices in the k-major table of the other paths.
__global__ void __launch_bounds__(TILE_BLOCK) k_tile_expand(Grid g, const int *__restrict__ cell_start, const Scalars *sc,
                                                            const unsigned short *__restrict__ nbrT, int cap,
                                                            const int *__restrict__ nbr_cnt, int *__restrict__ nbr, int npad)
{
    extern __shared__ __align__(16) unsigned char tile_smem[];
    int *slot_to_sorted = reinterpret_cast<int *>(tile_smem);  // one int per shell slot
    __shared__ TileTab T;
    tile_setup(T, g, cell_start, sc, blockIdx.x);
    for (int r = threadIdx.x >> 5; r < TILE_RUNS; r += TILE_WARPS)
        for (int e = threadIdx.x & 31; e < T.run_len[r]; e += 32) slot_to_sorted[T.run_loc[r] + e] = T.run_src[r] + e;
    __syncthreads();
    for (int ai = threadIdx.x >> 5; ai < T.n_own; ai += TILE_WARPS) {
        int oc, slot, gi;
        tile_locate(T, ai, oc, slot, gi);
        const int cnt = nbr_cnt[gi];
        for (int k = threadIdx.x & 31; k < cnt; k += 32)
            nbr[(size_t)k * npad + gi] = slot_to_sorted[nbrT[(size_t)gi * cap + (k & ~31) + tile_chunk_pos(k & 31)]];
    }
}

// ---- K3, tile form ----------------------------------------------------------------------------------------------------
// Pair phase: a warp takes FOUR brick atoms at a time, eight lanes each.  Lane (s, l) evaluates hits l, l + 8, l + 16, l + 24 of
// every 32-entry chunk of atom s's list per trip — four independent pair terms in flight, consecutive lanes on neighbouring
// slots — with the four 16-bit indices of the trip after next fetched as one 8-byte word before the current trip's arithmetic.  The
// partial forces of an atom are folded over its eight lanes with three xor-shuffles in a fixed order (deterministic): all
// the per-atom work — locate, count, fold, hand-over — is paid once per four atoms.
template <int UW>
__device__ __forceinline__ void tile_pair_phase(const TileTab &T, const double *__restrict__ sp,
                                                const unsigned short *__restrict__ nbrT, int cap,
                                                const int *__restrict__ nbr_cnt, const ForceConsts &fc, double *fst)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane >> 3, l8 = lane & 7;
    const LjConst c{};  // (no minimum image: the shell holds the images)
    const int n_own = T.n_own;
    // Software pipeline ACROSS groups: while a group's trips run, the next group's atom is located (shared memory only), its
    // list length and the four indices of its first trip are fetched — so no group starts with the chain
    // locate -> count -> indices -> positions exposed (each link a trip to L2 or HBM; it was a third of the first version's
    // time, profiles/r02_ncu_c5_force_tile_v2.txt) — and its whole list is pulled into L2 for the later trips.
    // Lane l8 of an atom takes hits l8, l8+8, l8+16, l8+24 of every 32-entry chunk of the atom's list — stored next to each other
    // (tile_chunk_pos): ONE 8-byte load per lane and trip.  Lists are padded to whole chunks with the nobody slot
    // (k_build_tile), chunks beyond an atom's list read as nobody.
    const unsigned int nobody2 = (unsigned int)T.n_shell * 0x10001u;
    const uint2 nobody = make_uint2(nobody2, nobody2);
    int n_slot = 0, n_gi = 0, n_cnt = 0;
    uint2 n_first[2] = {nobody, nobody};
    auto fetch_group = [&](int g0) {
        const int ai = g0 + sub;
        int oc;
        n_slot = 0; n_gi = 0; n_cnt = 0;
        if (ai < n_own) {
            tile_locate(T, ai, oc, n_slot, n_gi);
            n_cnt = nbr_cnt[n_gi];
        }
        const uint2 *__restrict__ lst = reinterpret_cast<const uint2 *>(nbrT + (size_t)n_gi * cap);  // (row 0 for lanes without an atom: never used)
        n_first[0] = lst[l8];      // unconditional: a row holds cap >= 64 entries
        n_first[1] = lst[8 + l8];
        const char *nl = reinterpret_cast<const char *>(lst);
        for (int b = (l8 + 1) * 128; b < 2 * cap; b += 8 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nl + b));
    };
    if (4 * warp < n_own) fetch_group(4 * warp);
    for (int g0 = 4 * warp; g0 < n_own; g0 += 4 * TILE_WARPS) {  // warp-uniform
        const int ai = g0 + sub;
        const int slot = n_slot, gi = n_gi, cnt = n_cnt;
        const int cntr = (cnt + 31) & ~31;
        // index registers two trips deep: a trip is shorter than a round trip to L2
        uint2 jn = 0 < cntr ? n_first[0] : nobody, jnn = 32 < cntr ? n_first[1] : nobody;
        if (g0 + 4 * TILE_WARPS < n_own) fetch_group(g0 + 4 * TILE_WARPS);
        const uint2 *__restrict__ lst = reinterpret_cast<const uint2 *>(nbrT + (size_t)gi * cap) + l8;
        const double xi = sp[3 * slot], yi = sp[3 * slot + 1], zi = sp[3 * slot + 2];
        const int kmax = __reduce_max_sync(0xffffffffu, cnt);
        PairAcc acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] = PairAcc{0.0, 0.0, 0.0, 0.0, 0.0};
        for (int k0 = 0; k0 < kmax; k0 += 32) {
            const uint2 jw = jn;
            jn = jnn;
            if (k0 + 64 < kmax) jnn = k0 + 64 < cntr ? lst[(k0 + 64) >> 2] : nobody;
            const int j[4] = {(int)(jw.x & 0xffffu), (int)(jw.x >> 16), (int)(jw.y & 0xffffu), (int)(jw.y >> 16)};
            double xj[4], yj[4], zj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) { const double *pj = sp + 3 * j[u]; xj[u] = pj[0]; yj[u] = pj[1]; zj[u] = pj[2]; }
#pragma unroll
            for (int u = 0; u < 4; ++u) pair_dense<false, UW>(acc[u], true, xj[u], yj[u], zj[u], xi, yi, zi, c, fc);
        }
        PairAcc f;
        f.fx = (acc[0].fx + acc[1].fx) + (acc[2].fx + acc[3].fx);
        f.fy = (acc[0].fy + acc[1].fy) + (acc[2].fy + acc[3].fy);
        f.fz = (acc[0].fz + acc[1].fz) + (acc[2].fz + acc[3].fz);
        f.w = UW >= 1 ? (acc[0].w + acc[1].w) + (acc[2].w + acc[3].w) : 0.0;
        f.u = UW >= 2 ? (acc[0].u + acc[1].u) + (acc[2].u + acc[3].u) : 0.0;
        // fixed-order fold over the atom's eight lanes
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            f.fx += __shfl_xor_sync(0xffffffffu, f.fx, o);
            f.fy += __shfl_xor_sync(0xffffffffu, f.fy, o);
            f.fz += __shfl_xor_sync(0xffffffffu, f.fz, o);
            if (UW >= 1) f.w += __shfl_xor_sync(0xffffffffu, f.w, o);
            if (UW >= 2) f.u += __shfl_xor_sync(0xffffffffu, f.u, o);
        }
        if (l8 == 0 && ai < n_own) {
            double *o = fst + (size_t)ai * 5;
            o[0] = f.fx; o[1] = f.fy; o[2] = f.fz; o[3] = f.u; o[4] = f.w;
        }
    }
}

__global__ void __launch_bounds__(TILE_BLOCK, 2)
    k_force_tile(Grid g, Arrays a, const int *__restrict__ cell_start, const unsigned short *__restrict__ nbrT, int cap,
                 const int *__restrict__ nbr_cnt, double *__restrict__ partials, Scalars *sc, const Params *__restrict__ pr,
                 int do_step, const ForceConsts fc, const int *__restrict__ brick_order, int sh_cap)
{
    // do_step bits: 1 = MD step (both half-kicks fused in), 4 = guarded (see k_force)
    if ((do_step & 4) && halted(sc)) return;  // uniform over the grid: nobody takes a ticket
    extern __shared__ __align__(16) unsigned char tile_smem[];
    double *sp = reinterpret_cast<double *>(tile_smem);  // 3 doubles per shell slot
    double *fst = sp + 3 * (size_t)sh_cap;                // per brick atom {fx, fy, fz, u, w}
    __shared__ TileTab T;
    tile_setup(T, g, cell_start, sc, brick_order[blockIdx.x]);
    tile_stage(T, a, sc, sp, true);
    const bool step = (do_step & 1) != 0;
    const bool store_state = !step || sc->steps_left <= 1;
    const bool nh = pr->th_kind == 2 || !step;
    // per-atom potential and virial enter nothing but the stored State and the S_U / S_W sums (see k_force)
    const bool need_u = store_state, need_w = store_state || pr->ba_kind != 0;
    if (need_u) tile_pair_phase<2>(T, sp, nbrT, cap, nbr_cnt, fc, fst);
    else if (need_w) tile_pair_phase<1>(T, sp, nbrT, cap, nbr_cnt, fc, fst);
    else tile_pair_phase<0>(T, sp, nbrT, cap, nbr_cnt, fc, fst);
    __syncthreads();
    // epilogue: one thread per brick atom — both half-kicks, K5 terms, stores (k_force's, same arithmetic)
    Sums s;
#pragma unroll
    for (int q = 0; q < NSUM; ++q) s.v[q] = 0.0;
    const double lambda = sc->lambda, c = fc.hc, mass = fc.mass;
    const double shift[3] = {sc->shift[0], sc->shift[1], sc->shift[2]};
    for (int ai = threadIdx.x; ai < T.n_own; ai += TILE_BLOCK) {
        int oc, slot, gi;
        tile_locate(T, ai, oc, slot, gi);
        const double *f = fst + (size_t)ai * 5;
        const double fx = f[0], fy = f[1], fz = f[2], fu = f[3], fw = f[4];
        double vx = a.vx[gi], vy = a.vy[gi], vz = a.vz[gi];
        if (step) {
            vx = __dadd_rn(__dmul_rn(vx, lambda), __dmul_rn(fx, c));  // v'' = lambda*u + F*c
            vy = __dadd_rn(__dmul_rn(vy, lambda), __dmul_rn(fy, c));
            vz = __dadd_rn(__dmul_rn(vz, lambda), __dmul_rn(fz, c));
        }
        const double wx = __dadd_rn(vx, __dmul_rn(fx, c)), wy = __dadd_rn(vy, __dmul_rn(fy, c)),
                     wz = __dadd_rn(vz, __dmul_rn(fz, c));  // u' = v'' + F*c
        s.v[0] += mass * vx; s.v[1] += mass * vy; s.v[2] += mass * vz;
        const double ax = vx - shift[0], ay = vy - shift[1], az = vz - shift[2];
        s.v[S_TH] += mass * (ax * ax + ay * ay + az * az);
        s.v[S_KE] += mass * (vx * vx + vy * vy + vz * vz);
        s.v[S_W] += fw;
        s.v[S_U] += fu;
        if (nh) {
            s.v[S_MU] += mass * wx; s.v[S_MU + 1] += mass * wy; s.v[S_MU + 2] += mass * wz;
            const double bx = wx - shift[0], by = wy - shift[1], bz = wz - shift[2];
            s.v[S_THU] += mass * (bx * bx + by * by + bz * bz);
        }
        s.v[S_MAX] = fmax(s.v[S_MAX], wx * wx + wy * wy + wz * wz);
        if (store_state) {
            a.fx[gi] = fx; a.fy[gi] = fy; a.fz[gi] = fz; a.u[gi] = fu; a.w[gi] = fw;
            if (step) { a.vx[gi] = vx; a.vy[gi] = vy; a.vz[gi] = vz; }
        } else {
            a.vx[gi] = wx; a.vy[gi] = wy; a.vz[gi] = wz;
        }
    }
    block_reduce<TILE_BLOCK>(s);
    grid_reduce_finalize<TILE_BLOCK>(s, partials, sc, pr, step ? FIN_STEP : 0, nullptr);
}

}  // namespace md

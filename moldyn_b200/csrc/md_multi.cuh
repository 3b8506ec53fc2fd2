// md_multi.cuh — States with several particle types.
//
// The reference indexes State.particles by type id (core/src/particle.rs:24-32) and, with T > 1 types, does things a
// single-type State never shows (DESIGN.md §1 has the written decision):
//   * update_force (potential.rs:168-176) loops `for t1 in 0..T { for t2 in t1..T {`: atoms of type t1 accumulate their
//     type-t2 partners, atoms of type t2 > t1 receive nothing from type t1.  MD_CROSS_REFERENCE reproduces that;
//     MD_CROSS_SYMMETRIC lets every atom accumulate every type (the symmetric type-pair table a force field means).
//   * Integrator::calculate (integrator.rs:18-27) calls calculate_myu / calculate_lambda once per type on the same
//     thermostat / barostat object: the LAST type's pressure and temperature decide myu and lambda for every type; the kicks
//     use the type's own mass (integrator.rs:29-30); barostat.update (integrator.rs:54-58) scales the box once per type.
//
// This is a correctness-first path: the device-resident State, cell sort and Verlet lists are the single-type ones (one list
// radius = the largest r_cut of the table + skin; EXACT lists are sorted by upload index = the reference's (type, index)
// order), the step is host-stepped (a few small kernels per step, one host look per step).  The single-type hot path does
// not go through any of this.
#pragma once
#include "md_common.cuh"
#include "md_force.cuh"

namespace md {

constexpr int MULTI_MAX_TYPES = 8;
constexpr int MULTI_BLOCK = 256;
// per-type sums, pass A: Σvx Σvy Σvz Σ|v|² ΣU ΣW ; pass B (needs the type's COM velocity): Σdvx² Σdvy² Σdvz²
constexpr int MULTI_NA = 6, MULTI_NB = 3;

struct MultiTable {
    int T, symmetric;
    int start[MULTI_MAX_TYPES + 1];  // upload index of the first atom of every type (atoms are uploaded type by type)
    double mass[MULTI_MAX_TYPES];
    double hc[MULTI_MAX_TYPES];      // dt / (2.0 * mass)   integrator.rs:30
    // PotentialsDatabase::get_potential(t1, t2), entry [t1 * T + t2] (symmetric: keyed (min, max), potential.rs:147-155)
    double sigma[MULTI_MAX_TYPES * MULTI_MAX_TYPES], eps4[MULTI_MAX_TYPES * MULTI_MAX_TYPES],
        eps24[MULTI_MAX_TYPES * MULTI_MAX_TYPES], r_cut[MULTI_MAX_TYPES * MULTI_MAX_TYPES],
        u_cut[MULTI_MAX_TYPES * MULTI_MAX_TYPES];
};

struct MultiTypeSums {
    double a[MULTI_NA];
    double b[MULTI_NB];
    // derived by k_multi_controls
    double vcom[3], kinetic, thermal, potential, temperature, pressure;
};

struct MultiWork {
    MultiTypeSums type[MULTI_MAX_TYPES];
    unsigned long long vmax2_bits;  // max |v|² of the last drift, as the bits of a non-negative double (ordered like integers)
};

__device__ __forceinline__ int multi_type_of(int id, const int *start, int T)
{
    int t = 0;
    while (t + 1 < T && id >= start[t + 1]) ++t;
    return t;
}

// deterministic block sum of K values per thread: lane tree, then warp 0 adds the warps' results in order
template <int K>
__device__ __forceinline__ void multi_block_sum(double (&v)[K], double *out /* K */)
{
    __shared__ double red[K][MULTI_BLOCK / 32];
#pragma unroll
    for (int q = 0; q < K; ++q) {
        double x = v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if ((threadIdx.x & 31) == 0) red[q][threadIdx.x >> 5] = x;
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double s = 0.0;
        for (int w = 0; w < MULTI_BLOCK / 32; ++w) s += red[threadIdx.x][w];
        out[threadIdx.x] = s;
    }
    __syncthreads();
}

// Per-type sums over the cell-sorted planes: blockIdx.y = type, every block row walks all atoms and keeps its type's.
// pass 0: MULTI_NA sums; pass 1: MULTI_NB sums around the type's COM velocity (written by k_multi_fold after pass 0).
__global__ void __launch_bounds__(MULTI_BLOCK) k_multi_sums(int n, Arrays a, const MultiTable *__restrict__ tab,
                                                            const MultiWork *__restrict__ work, int pass,
                                                            double *__restrict__ partials /* [T][gridDim.x][6] */)
{
    const int t = blockIdx.y;
    const int lo = tab->start[t], hi = tab->start[t + 1];
    double v[MULTI_NA] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
    const double cx = pass ? work->type[t].vcom[0] : 0.0, cy = pass ? work->type[t].vcom[1] : 0.0,
                 cz = pass ? work->type[t].vcom[2] : 0.0;
    for (int i = blockIdx.x * MULTI_BLOCK + threadIdx.x; i < n; i += gridDim.x * MULTI_BLOCK) {
        const int id = a.id[i];
        if (id < lo || id >= hi) continue;
        const double vx = a.vx[i], vy = a.vy[i], vz = a.vz[i];
        if (pass == 0) {
            v[0] += vx; v[1] += vy; v[2] += vz;
            v[3] += (vx * vx + vy * vy) + vz * vz;
            v[4] += a.u[i];
            v[5] += a.w[i];
        } else {
            const double dx = vx - cx, dy = vy - cy, dz = vz - cz;
            v[0] += dx * dx; v[1] += dy * dy; v[2] += dz * dz;
        }
    }
    multi_block_sum<MULTI_NA>(v, partials + ((size_t)t * gridDim.x + blockIdx.x) * MULTI_NA);
}

// One block per type folds the partials in block order.  After pass 0 it also forms the type's COM velocity
// (get_center_of_mass_velocity, mod.rs:12-25: Σ(v·m) / Σm — one mass per type, so Σv / count up to rounding).
__global__ void __launch_bounds__(32) k_multi_fold(int nblocks, const double *__restrict__ partials,
                                                   const MultiTable *__restrict__ tab, MultiWork *work, int pass)
{
    const int t = blockIdx.x, q = threadIdx.x;
    if (q >= MULTI_NA) return;
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partials[((size_t)t * nblocks + b) * MULTI_NA + q];
    MultiTypeSums &ts = work->type[t];
    if (pass == 0) {
        ts.a[q] = s;
        if (q < 3) {
            const double cnt = (double)(tab->start[t + 1] - tab->start[t]), m = tab->mass[t];
            ts.vcom[q] = (s * m) / (cnt * m);
        }
    } else if (q < MULTI_NB) {
        ts.b[q] = s;
    }
}

// The reference's per-type macro parameters, then integrator.rs:18-27: calculate_myu / calculate_lambda type by type on the
// same object — the last type's values stay (Nose-Hoover is not offered for T > 1: its psi would need a second reduction
// between the kick and the scaling of every type, thermostat.rs:47-65).
__global__ void k_multi_controls(const MultiTable *__restrict__ tab, MultiWork *work, Scalars *sc, const Params *__restrict__ pr,
                                 int apply)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int T = tab->T;
    const double volume = sc->box[0] * sc->box[1] * sc->box[2];
    double lambda = 1.0, mu = 1.0;
    for (int t = 0; t < T; ++t) {
        MultiTypeSums &ts = work->type[t];
        const double m = tab->mass[t];
        const double cnt = (double)(tab->start[t + 1] - tab->start[t]);
        ts.kinetic = m * ts.a[3] / 2.0;                                  // energy.rs:4-22
        ts.thermal = m * ((ts.b[0] + ts.b[1]) + ts.b[2]) / 2.0;          // energy.rs:8-37
        ts.potential = ts.a[4] / 2.0;                                    // energy.rs:40-49
        ts.temperature = (2.0 * ts.thermal) / (3.0 * cnt * K_B) * 100.0;  // temperature.rs:4-7
        const double r1 = m * ts.b[0] + m * ts.b[1] + m * ts.b[2], r2 = -ts.a[5];
        ts.pressure = (r1 + r2 * 0.5) / volume / 3.0;                    // pressure.rs:5-20
        if (apply && pr->ba_kind == 1) mu = cbrt(1.0 + pr->dt * pr->ba_beta / pr->ba_tau * (ts.pressure - pr->ba_target));
        if (apply && pr->th_kind == 1) lambda = sqrt(1.0 + pr->dt / pr->th_tau * (pr->th_target / ts.temperature - 1.0));
    }
    if (apply) {
        if (!isfinite(lambda) || !isfinite(mu) || !(mu > 0.0)) sc->error = 7;  // MD_ERR_NONFINITE
        sc->lambda = lambda; sc->mu = mu;
        sc->lambda_last = lambda; sc->mu_last = mu;
        sc->temperature = work->type[T - 1].temperature;
        sc->pressure = work->type[T - 1].pressure;
        work->vmax2_bits = 0ull;
    }
}

// integrator.rs:28-45 for every atom: v = v + F·(dt/2m_type); v *= lambda; x += v·dt; single-shift wrap.  Operation by
// operation like the reference (no contraction) — the same code serves the EXACT and the FAST mode.
__global__ void __launch_bounds__(MULTI_BLOCK) k_multi_kick_drift(int n, Arrays a, const MultiTable *__restrict__ tab, Scalars *sc,
                                                                  const Params *__restrict__ pr, MultiWork *work)
{
    const int i = blockIdx.x * MULTI_BLOCK + threadIdx.x;
    double w2 = 0.0;
    if (i < n) {
        const int t = multi_type_of(a.id[i], tab->start, tab->T);
        const double c = tab->hc[t], lambda = sc->lambda, dt = pr->dt;
        const bool scale = pr->th_kind != 0;
        double vx = __dadd_rn(a.vx[i], __dmul_rn(a.fx[i], c)), vy = __dadd_rn(a.vy[i], __dmul_rn(a.fy[i], c)),
               vz = __dadd_rn(a.vz[i], __dmul_rn(a.fz[i], c));
        if (scale) { vx = __dmul_rn(vx, lambda); vy = __dmul_rn(vy, lambda); vz = __dmul_rn(vz, lambda); }
        double x = __dadd_rn(a.x[i], __dmul_rn(vx, dt)), y = __dadd_rn(a.y[i], __dmul_rn(vy, dt)),
               z = __dadd_rn(a.z[i], __dmul_rn(vz, dt));
        const double Lx = sc->box[0], Ly = sc->box[1], Lz = sc->box[2];
        if (x < 0.0) x = __dadd_rn(x, Lx); else if (x >= Lx) x = __dsub_rn(x, Lx);
        if (y < 0.0) y = __dadd_rn(y, Ly); else if (y >= Ly) y = __dsub_rn(y, Ly);
        if (z < 0.0) z = __dadd_rn(z, Lz); else if (z >= Lz) z = __dsub_rn(z, Lz);
        a.vx[i] = vx; a.vy[i] = vy; a.vz[i] = vz;
        a.x[i] = x; a.y[i] = y; a.z[i] = z;
        w2 = vx * vx + vy * vy + vz * vz;
        if (!(w2 == w2)) w2 = __longlong_as_double(0x7ff0000000000000ll);  // NaN → +inf: the host sees a non-finite bound
    }
    // max over the grid: non-negative doubles order like their bit patterns
    unsigned long long b = (unsigned long long)__double_as_longlong(w2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, b, o);
        b = other > b ? other : b;
    }
    if ((threadIdx.x & 31) == 0 && b) atomicMax(&work->vmax2_bits, b);
}

// update_force for every atom from its Verlet list (built with the largest r_cut of the table), the pair's own potential
// from the type-pair table, then — inside a step — the second half-kick with the type's mass (integrator.rs:47-53).
// EXACT: pair_exact() in list order = ascending upload index = the reference's (t2, j) order.  FAST: the same formula with
// the compiler free to contract (r-form: the per-pair constants differ, the r²-form's precomputed powers would be a table of
// their own), cell order.
template <bool EXACT>
__global__ void __launch_bounds__(MULTI_BLOCK) k_multi_force(int n, int npad, Arrays a, const int *__restrict__ nbr,
                                                             const int *__restrict__ nbr_cnt,
                                                             const MultiTable *__restrict__ tab, const Scalars *__restrict__ sc,
                                                             int kick)
{
    __shared__ MultiTable st;
    for (int k = threadIdx.x; k < (int)(sizeof(MultiTable) / sizeof(int)); k += MULTI_BLOCK)
        reinterpret_cast<int *>(&st)[k] = reinterpret_cast<const int *>(tab)[k];
    __syncthreads();
    const int i = blockIdx.x * MULTI_BLOCK + threadIdx.x;
    if (i >= n) return;
    LjConst c;
    c.Lx = sc->box[0]; c.Ly = sc->box[1]; c.Lz = sc->box[2];
    c.hx = c.Lx / 2.0; c.hy = c.Ly / 2.0; c.hz = c.Lz / 2.0;
    c.hxi = __double2hiint(c.hx); c.hyi = __double2hiint(c.hy); c.hzi = __double2hiint(c.hz);
    const int T = st.T;
    const int ti = multi_type_of(a.id[i], st.start, T);
    const double xi = a.x[i], yi = a.y[i], zi = a.z[i];
    const int cnt = nbr_cnt[i];
    PairAcc f = {0.0, 0.0, 0.0, 0.0, 0.0};
    for (int k = 0; k < cnt; ++k) {
        const int j = nbr[(size_t)k * npad + i];
        const int tj = multi_type_of(a.id[j], st.start, T);
        if (!st.symmetric && tj < ti) continue;  // potential.rs:169: `for particle_type2 in particle_type1..`
        const int e = ti * T + tj;
        ForceConsts fc{};
        fc.sigma = st.sigma[e]; fc.eps4 = st.eps4[e]; fc.eps24 = st.eps24[e]; fc.r_cut = st.r_cut[e]; fc.u_cut = st.u_cut[e];
        if (EXACT) {
            pair_exact(f, a.x[j], a.y[j], a.z[j], xi, yi, zi, c, fc);
        } else {
            const double rx = min_image(a.x[j] - xi, c.Lx, c.hx), ry = min_image(a.y[j] - yi, c.Ly, c.hy),
                         rz = min_image(a.z[j] - zi, c.Lz, c.hz);
            const double r2 = rx * rx + ry * ry + rz * rz;
            if (r2 > fc.r_cut * fc.r_cut) continue;
            const double inv = 1.0 / r2, s2 = fc.sigma * fc.sigma * inv, s6 = s2 * s2 * s2, s12 = s6 * s6;
            const double fr = fc.eps24 * inv * (s6 - 2.0 * s12);  // F / r
            f.fx += fr * rx; f.fy += fr * ry; f.fz += fr * rz;
            f.u += fc.eps4 * (s12 - s6) - fc.u_cut;
            f.w += fr * r2;
        }
    }
    a.fx[i] = f.fx; a.fy[i] = f.fy; a.fz[i] = f.fz;
    a.u[i] = f.u; a.w[i] = f.w;
    if (kick) {
        const double hc = st.hc[ti];
        a.vx[i] = __dadd_rn(a.vx[i], __dmul_rn(f.fx, hc));
        a.vy[i] = __dadd_rn(a.vy[i], __dmul_rn(f.fy, hc));
        a.vz[i] = __dadd_rn(a.vz[i], __dmul_rn(f.fz, hc));
    }
}

// barostat.update once per type (integrator.rs:54-58, barostat.rs:39-49): every call multiplies the box — myu^T in all —
// and the positions of its own type once.
__global__ void __launch_bounds__(MULTI_BLOCK) k_multi_scale(int n, Arrays a, const MultiTable *__restrict__ tab, Scalars *sc)
{
    const int i = blockIdx.x * MULTI_BLOCK + threadIdx.x;
    const double mu = sc->mu;
    if (i < n) {
        a.x[i] = __dmul_rn(a.x[i], mu); a.y[i] = __dmul_rn(a.y[i], mu); a.z[i] = __dmul_rn(a.z[i], mu);
    }
}

__global__ void k_multi_scale_box(const MultiTable *__restrict__ tab, Scalars *sc)
{
    const double mu = sc->mu;
    for (int t = 0; t < tab->T; ++t) {
        sc->box[0] = __dmul_rn(sc->box[0], mu); sc->box[1] = __dmul_rn(sc->box[1], mu); sc->box[2] = __dmul_rn(sc->box[2], mu);
    }
}

}  // namespace md

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

__device__ __forceinline__ void kick_drift_tail(int i, Arrays a, double lambda, double mup, double Lx, double Ly, double Lz,
                                                bool half, const Params *__restrict__ pr, bool write_q4)
{
    const double c = pr->half_dt_m, dt = pr->dt;
    double ux = a.vx[i], uy = a.vy[i], uz = a.vz[i];
    if (!half) {
        ux = __dadd_rn(ux, __dmul_rn(a.fx[i], c)); uy = __dadd_rn(uy, __dmul_rn(a.fy[i], c));
        uz = __dadd_rn(uz, __dmul_rn(a.fz[i], c));
        a.vx[i] = ux; a.vy[i] = uy; a.vz[i] = uz;
    }
    double x = a.x[i], y = a.y[i], z = a.z[i];
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

__global__ void __launch_bounds__(256, 4) k_kick_drift(int n, Arrays a, Scalars *sc, const Params *__restrict__ pr,
                                                    int guarded, int write_q4, const HaloPush h)
{
    // guarded: return at once when the loop is halted (speculatively enqueued steps)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (guarded && halted(sc)) return;
    const double lambda = sc->lambda, mup = sc->mu_pending;
    const double Lx = sc->box[0], Ly = sc->box[1], Lz = sc->box[2];
    const bool half = sc->vel_is_half != 0;
    // block-uniform: does this block hold face atoms?  (512 atoms per block)
    const int b_lo = blockIdx.x * 512, b_hi = b_lo + 512;
    const bool pushes = (h.m[0] | h.m[1]) != 0 && (b_lo < h.m[0] || b_hi > n - h.m[1]);
    if (2 * t < n) {
        if (2 * t + 1 >= n) {  // odd tail: one atom, scalar accesses (the slot after it may belong to a ghost atom)
            kick_drift_tail(2 * t, a, lambda, mup, Lx, Ly, Lz, half, pr, write_q4 != 0);
            if (pushes) push_atom(h, 2 * t, n, a.x[2 * t], a.y[2 * t], a.z[2 * t]);
        } else {
            const double c = pr->half_dt_m, dt = pr->dt;
            double2 x = reinterpret_cast<double2 *>(a.x)[t], y = reinterpret_cast<double2 *>(a.y)[t],
                    z = reinterpret_cast<double2 *>(a.z)[t];
            double2 ux = reinterpret_cast<double2 *>(a.vx)[t], uy = reinterpret_cast<double2 *>(a.vy)[t],
                    uz = reinterpret_cast<double2 *>(a.vz)[t];
            if (!half) {
                const double2 fx = reinterpret_cast<const double2 *>(a.fx)[t], fy = reinterpret_cast<const double2 *>(a.fy)[t],
                              fz = reinterpret_cast<const double2 *>(a.fz)[t];
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

// ---- mbarrier / TMA bulk-copy helpers (used by the tile kernels of dense systems, md_tile.cuh) ----------------------
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

// ---- measurement aid: the FP64 pipe's DFMA rate (the roofline denominator of the dense force kernel) -----------------
// 8 independent fused multiply-add chains per thread, no memory traffic: 2 * 8 * iters flop per thread.
__global__ void __launch_bounds__(256) k_fp64_peak(int iters, double seed, double *__restrict__ sink)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0,
           a7 = a0 + 7.0;
    const double m = 0.999999, c = 1e-9;
#pragma unroll 4
    for (int k = 0; k < iters; ++k) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    const double r = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (r == 12345.678) sink[0] = r;  // keeps the chains alive
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

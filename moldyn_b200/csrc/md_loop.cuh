// md_loop.cuh — the persistent step loop of dilute systems: ONE cooperative kernel runs MD steps until the neighbour list
// has to be rebuilt (or the batch ends), with two grid-wide synchronisations per step instead of two kernel boundaries.
// Part of md_kernels.cuh (included from there, in order; one translation unit).
#pragma once

namespace md {

// Why.  Round-1 measurements: at 10^6 atoms the two-kernel step spends 21 of its 30 us in a force kernel
// that is bound by gather latency, not bandwidth — every warp walks the list path although 60-80 % of the atoms of the gas
// have no listed partner — and below ~10^5 atoms per GPU (the 8-GPU slabs, C1, C2) the step is nothing but fixed latencies:
// two launches, a 592-slot fold, a mailbox exchange.  This kernel restructures the step (integrator.rs:14-59) as
//
//   phase A  (all atoms, two per thread, 128-bit accesses)  first half-kick on the first step of a batch, lambda-scale,
//            pending barostat scale, drift, wrap — k_kick_drift's arithmetic — and, for the atoms WITHOUT listed partners,
//            the rest of the step as well: F = 0, so v'' = u' = lambda*u is final, its K5 terms are summed here and its
//            velocity is stored once.  Multi-GPU: a few "face blocks" handle the atoms the neighbours see as ghosts first,
//            store them into the neighbours' planes (NVLink), fence, and raise the neighbours' halo flags — while the rest
//            of the grid is still drifting.
//   barrier  every drifted position is visible to the whole grid.
//   phase B  only the atoms WITH listed partners, one per lane through the index list compacted at the last rebuild (every
//            lane has gather work): pair forces, both half-kicks, K5 terms.  Multi-GPU: interior atoms first; the atoms
//            whose lists hold ghosts wait for the neighbours' flags only when they get there.
//   tail     block sums -> ticket -> the last block folds, exchanges the rank sums through the peer mailboxes, finalizes
//            (T, P, lambda, myu, rebuild decision) and release-stores the step's sequence number; the other blocks wait for
//            it and start the next step.
//
// Per-atom arithmetic is the same finish_atom()/drift_one()/pair term as in the two-kernel path.  Everything that another
// block (or GPU) wrote during the launch is read with ld.global.cg (L2): the L1 is not coherent across SMs.
constexpr int LOOP_BLOCK = 512;
constexpr int LOOP_FACE_BLOCKS = 8;  // blocks that drift and push the face atoms first (multi-GPU)

struct LoopArgs {
    int n;                  // owned atoms
    int npad, cap;          // row stride and capacity of the neighbour table
    Arrays a;
    const int *nbr, *nbr_cnt;
    const int *act_idx;     // atoms with >= 1 listed partner: [0, n_act[0]) interior, then n_act[1] whose lists hold ghosts
    const int *n_act;       // (device) the two counts
    double *partials;
    Scalars *sc;
    const Params *pr;
    const Peers *peers;     // NULL on one GPU
    HaloPush h;             // where the face atoms land in the neighbours' planes (m = {0, 0} on one GPU)
    long long max_steps;    // steps this launch may run (host-stepped loop: 1)
    ForceConsts fc;
};

struct LoopCtl {
    double lambda, mup, Lx, Ly, Lz, shift[3];
    long long steps_left;
    unsigned long long fin_seq, epoch;
    int halted, half;
};

template <bool EXACT>
__global__ void __launch_bounds__(LOOP_BLOCK, 1) k_md_loop(const LoopArgs A)
{
    extern __shared__ __align__(16) unsigned char loop_smem[];
    SumsSmemT<LOOP_BLOCK> &ss = *reinterpret_cast<SumsSmemT<LOOP_BLOCK> *>(loop_smem);
    __shared__ LoopCtl ctl;
    const int tid = threadIdx.x, bid = blockIdx.x, nb = gridDim.x;
    Scalars *sc = A.sc;
    const Params *__restrict__ pr = A.pr;
    const Arrays &a = A.a;
    const ForceConsts &fc = A.fc;
    const int n = A.n;
    const int npairs = (n + 1) >> 1;
    const bool multi = A.peers != nullptr;
    // face pairs: every pair that holds an atom of the prefix [0, m0) or of the suffix [n - m1, n)
    const int pl = multi ? min((A.h.m[0] + 1) >> 1, npairs) : 0;
    const int pr0 = multi ? max(min((n - A.h.m[1]) >> 1, npairs), pl) : npairs;
    const int nface = pl + (npairs - pr0);
    const int nfb = min(LOOP_FACE_BLOCKS, nb);
    const bool nh = pr->th_kind == 2;
    const double dt = pr->dt, hc = fc.hc, mass = fc.mass;
    const PairAcc zero = {0.0, 0.0, 0.0, 0.0, 0.0};
    __shared__ unsigned long long t_acc[5];  // block 0's phase clocks + step count (thread 0)
    if (tid == 0) { t_acc[0] = t_acc[1] = t_acc[2] = t_acc[3] = t_acc[4] = 0ull; }

    for (long long it = 0; it < A.max_steps; ++it) {
        // ---- step controls: written by the last finalize before its release, uniform over the grid ----------------------
        if (tid == 0) {
            ctl.steps_left = __ldcg(&sc->steps_left);
            ctl.halted = (__ldcg(&sc->need_rebuild) != 0 || __ldcg(&sc->error) != 0 || ctl.steps_left <= 0) ? 1 : 0;
            ctl.lambda = __ldcg(&sc->lambda); ctl.mup = __ldcg(&sc->mu_pending);
            ctl.Lx = __ldcg(&sc->box[0]); ctl.Ly = __ldcg(&sc->box[1]); ctl.Lz = __ldcg(&sc->box[2]);
            ctl.shift[0] = __ldcg(&sc->shift[0]); ctl.shift[1] = __ldcg(&sc->shift[1]); ctl.shift[2] = __ldcg(&sc->shift[2]);
            ctl.half = __ldcg(&sc->vel_is_half);
            ctl.fin_seq = __ldcg(&sc->fin_seq);
            ctl.epoch = __ldcg(&sc->epoch);
        }
        __syncthreads();
        if (ctl.halted) break;
        unsigned long long tq = tid == 0 ? gtime() : 0ull;
        MD_TRACE(bid == 0 && tid == 0, 0);
        const double lambda = ctl.lambda, mup = ctl.mup, Lx = ctl.Lx, Ly = ctl.Ly, Lz = ctl.Lz;
        const double *shift = ctl.shift;
        const bool half = ctl.half != 0;
        const bool store_state = ctl.steps_left <= 1;  // last step of the batch: the resident State must be complete
#pragma unroll
        for (int q = 0; q < NSUM; ++q) ss.v[q][tid] = 0.0;

        // ---- phase A ----------------------------------------------------------------------------------------------------
        auto drift_pair = [&](int t, bool face) {
            const int i0 = 2 * t;
            if (i0 + 1 >= n) {  // odd tail: one atom, scalar accesses (the slot after it may belong to a ghost atom)
                double ux = __ldcg(a.vx + i0), uy = __ldcg(a.vy + i0), uz = __ldcg(a.vz + i0);
                if (!half) {
                    ux = __dadd_rn(ux, __dmul_rn(__ldcg(a.fx + i0), hc)); uy = __dadd_rn(uy, __dmul_rn(__ldcg(a.fy + i0), hc));
                    uz = __dadd_rn(uz, __dmul_rn(__ldcg(a.fz + i0), hc));
                }
                double x = __ldcg(a.x + i0), y = __ldcg(a.y + i0), z = __ldcg(a.z + i0);
                drift_one(x, ux, lambda, mup, dt, Lx); drift_one(y, uy, lambda, mup, dt, Ly); drift_one(z, uz, lambda, mup, dt, Lz);
                a.x[i0] = x; a.y[i0] = y; a.z[i0] = z;
                if (face) push_atom(A.h, i0, n, x, y, z);
                if (A.nbr_cnt[i0] == 0) {
                    double wx, wy, wz;
                    finish_atom(ss, zero, ux, uy, uz, true, lambda, hc, mass, shift, wx, wy, wz, nh);
                    ux = wx; uy = wy; uz = wz;
                    if (store_state) { a.fx[i0] = 0.0; a.fy[i0] = 0.0; a.fz[i0] = 0.0; a.u[i0] = 0.0; a.w[i0] = 0.0; }
                }
                a.vx[i0] = ux; a.vy[i0] = uy; a.vz[i0] = uz;
                return;
            }
            double2 x = __ldcg(reinterpret_cast<const double2 *>(a.x) + t), y = __ldcg(reinterpret_cast<const double2 *>(a.y) + t),
                    z = __ldcg(reinterpret_cast<const double2 *>(a.z) + t);
            double2 ux = __ldcg(reinterpret_cast<const double2 *>(a.vx) + t), uy = __ldcg(reinterpret_cast<const double2 *>(a.vy) + t),
                    uz = __ldcg(reinterpret_cast<const double2 *>(a.vz) + t);
            const int2 C = reinterpret_cast<const int2 *>(A.nbr_cnt)[t];
            if (!half) {  // integrator.rs:28-34 on the first step of a batch
                const double2 fx = __ldcg(reinterpret_cast<const double2 *>(a.fx) + t),
                              fy = __ldcg(reinterpret_cast<const double2 *>(a.fy) + t),
                              fz = __ldcg(reinterpret_cast<const double2 *>(a.fz) + t);
                ux.x = __dadd_rn(ux.x, __dmul_rn(fx.x, hc)); ux.y = __dadd_rn(ux.y, __dmul_rn(fx.y, hc));
                uy.x = __dadd_rn(uy.x, __dmul_rn(fy.x, hc)); uy.y = __dadd_rn(uy.y, __dmul_rn(fy.y, hc));
                uz.x = __dadd_rn(uz.x, __dmul_rn(fz.x, hc)); uz.y = __dadd_rn(uz.y, __dmul_rn(fz.y, hc));
            }
            drift_one(x.x, ux.x, lambda, mup, dt, Lx); drift_one(x.y, ux.y, lambda, mup, dt, Lx);
            drift_one(y.x, uy.x, lambda, mup, dt, Ly); drift_one(y.y, uy.y, lambda, mup, dt, Ly);
            drift_one(z.x, uz.x, lambda, mup, dt, Lz); drift_one(z.y, uz.y, lambda, mup, dt, Lz);
            reinterpret_cast<double2 *>(a.x)[t] = x; reinterpret_cast<double2 *>(a.y)[t] = y;
            reinterpret_cast<double2 *>(a.z)[t] = z;
            if (face) {
                push_atom(A.h, i0, n, x.x, y.x, z.x);
                push_atom(A.h, i0 + 1, n, x.y, y.y, z.y);
            }
            // atoms without listed partners: F = 0, the step ends here (v'' = u' = lambda*u); the others keep u for phase B
            const bool s0 = C.x == 0, s1 = C.y == 0;
            if (s0) {
                double wx, wy, wz;
                finish_atom(ss, zero, ux.x, uy.x, uz.x, true, lambda, hc, mass, shift, wx, wy, wz, nh);
                ux.x = wx; uy.x = wy; uz.x = wz;
            }
            if (s1) {
                double wx, wy, wz;
                finish_atom(ss, zero, ux.y, uy.y, uz.y, true, lambda, hc, mass, shift, wx, wy, wz, nh);
                ux.y = wx; uy.y = wy; uz.y = wz;
            }
            if (!half || s0 || s1) {
                reinterpret_cast<double2 *>(a.vx)[t] = ux; reinterpret_cast<double2 *>(a.vy)[t] = uy;
                reinterpret_cast<double2 *>(a.vz)[t] = uz;
            }
            if (store_state) {
                if (s0 && s1) {
                    const double2 z2 = make_double2(0.0, 0.0);
                    reinterpret_cast<double2 *>(a.fx)[t] = z2; reinterpret_cast<double2 *>(a.fy)[t] = z2;
                    reinterpret_cast<double2 *>(a.fz)[t] = z2; reinterpret_cast<double2 *>(a.u)[t] = z2;
                    reinterpret_cast<double2 *>(a.w)[t] = z2;
                } else {
                    if (s0) { a.fx[i0] = 0.0; a.fy[i0] = 0.0; a.fz[i0] = 0.0; a.u[i0] = 0.0; a.w[i0] = 0.0; }
                    if (s1) { a.fx[i0 + 1] = 0.0; a.fy[i0 + 1] = 0.0; a.fz[i0 + 1] = 0.0; a.u[i0 + 1] = 0.0; a.w[i0 + 1] = 0.0; }
                }
            }
        };
        if (multi && bid < nfb) {
            // face atoms first; once every face block's stores are fenced system-wide, the last of them raises the flags
            for (int q = bid * LOOP_BLOCK + tid; q < nface; q += nfb * LOOP_BLOCK) drift_pair(q < pl ? q : pr0 + (q - pl), true);
            __syncthreads();
            if (tid == 0) {
                __threadfence_system();
                const unsigned int old = atomicAdd(&sc->face_arrive[0], 1u);
                if (old == (unsigned int)nfb - 1u) {
                    __threadfence_system();
                    st_release_sys(&A.peers->mail[A.peers->left]->halo_seq[1], ctl.epoch + 1);   // we are the left neighbour's right side
                    st_release_sys(&A.peers->mail[A.peers->right]->halo_seq[0], ctl.epoch + 1);
                }
            }
        }
        for (int r = bid * LOOP_BLOCK + tid; r < pr0 - pl; r += nb * LOOP_BLOCK) drift_pair(pl + r, false);
        if (tid == 0) { const unsigned long long t = gtime(); t_acc[0] += t - tq; tq = t; }
        MD_TRACE(bid == 0 && tid == 0, 1);

        // ---- mid-step barrier: all drifted positions (and first-step half-kicks) are visible grid-wide ------------------
        __syncthreads();
        if (tid == 0) {
            // release: this block's drifted positions (the bar.sync above makes the release cumulative over its threads);
            // the acquiring loads below also invalidate this SM's L1
            atom_add_release_gpu(&sc->bar_arrive, 1u);
            while (ld_acquire_gpu(&sc->bar_arrive) < (unsigned int)nb) { }
        }
        __syncthreads();
        if (tid == 0) { const unsigned long long t = gtime(); t_acc[1] += t - tq; tq = t; }
        MD_TRACE(bid == 0 && tid == 0, 2);

        // ---- phase B ----------------------------------------------------------------------------------------------------
        LjConst c;
        c.Lx = Lx; c.Ly = Ly; c.Lz = Lz;
        c.hx = Lx / 2.0; c.hy = Ly / 2.0; c.hz = Lz / 2.0;
        c.hxi = __double2hiint(c.hx); c.hyi = __double2hiint(c.hy); c.hzi = __double2hiint(c.hz);
        auto force_atom = [&](int i) {
            const double xi = __ldcg(a.x + i), yi = __ldcg(a.y + i), zi = __ldcg(a.z + i);
            double vx = __ldcg(a.vx + i), vy = __ldcg(a.vy + i), vz = __ldcg(a.vz + i);
            const int cnt = A.nbr_cnt[i];
            PairAcc f = zero;
            // partners in groups of four: the four indices are fetched together, then the twelve coordinates, then the pair
            // terms — two dependent memory round trips per group instead of two per partner
            for (int k0 = 0; k0 < cnt; k0 += 4) {
                int j[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) j[u] = k0 + u < cnt ? A.nbr[(size_t)(k0 + u) * A.npad + i] : i;
                double xj[4], yj[4], zj[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) { xj[u] = __ldcg(a.x + j[u]); yj[u] = __ldcg(a.y + j[u]); zj[u] = __ldcg(a.z + j[u]); }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (k0 + u < cnt) {
                        if (EXACT) pair_exact(f, xj[u], yj[u], zj[u], xi, yi, zi, c, fc);
                        else pair_fast_branchy(f, true, xj[u], yj[u], zj[u], xi, yi, zi, c, fc);
                    }
                }
            }
            double wx, wy, wz;
            finish_atom(ss, f, vx, vy, vz, true, lambda, hc, mass, shift, wx, wy, wz, nh);
            if (store_state) {
                a.fx[i] = f.fx; a.fy[i] = f.fy; a.fz[i] = f.fz; a.u[i] = f.u; a.w[i] = f.w;
                a.vx[i] = vx; a.vy[i] = vy; a.vz[i] = vz;
            } else {
                a.vx[i] = wx; a.vy[i] = wy; a.vz[i] = wz;
            }
        };
        const int n_int = A.n_act[0], n_bnd = A.n_act[1];
        for (int k = bid * LOOP_BLOCK + tid; k < n_int; k += nb * LOOP_BLOCK) force_atom(A.act_idx[k]);
        if (multi) {
            // the neighbours' ghosts of this step (their face blocks fenced the stores before raising the flag)
            __shared__ int halo_late;
            if (tid == 0) {
                const Mail *own = A.peers->mail[A.peers->rank];
                const unsigned long long tw = gtime();
                const unsigned long long seq = ctl.epoch + 1;
                halo_late = !(wait_seq(&own->halo_seq[0], seq) && wait_seq(&own->halo_seq[1], seq));
                if (bid == 0) sc->wait_halo_ns += gtime() - tw;
            }
            __syncthreads();
            if (halo_late && tid == 0) atomicExch(&sc->error, 3);
            for (int k = bid * LOOP_BLOCK + tid; k < n_bnd; k += nb * LOOP_BLOCK) force_atom(A.act_idx[n_int + k]);
        }
        if (tid == 0) { const unsigned long long t = gtime(); t_acc[2] += t - tq; tq = t; }
        MD_TRACE(bid == 0 && tid == 0, 3);

        // ---- tail: block sums, ticket, last block folds + finalizes + releases -------------------------------------------
        Sums s;
#pragma unroll
        for (int q = 0; q < NSUM; ++q) s.v[q] = ss.v[q][tid];
        block_reduce<LOOP_BLOCK>(s);
        MD_TRACE(bid == 0 && tid == 0, 4);
        const bool last = publish_and_ticket<LOOP_BLOCK>(s, A.partials, sc);
        MD_TRACE(bid == 0 && tid == 0, 5);
        if (last)
            last_block_finalize<LOOP_BLOCK>(A.partials, sc, pr, FIN_STEP | (multi ? FIN_P2P : 0), A.peers);
        if (tid == 0) {
            const unsigned long long fin_seq0 = ctl.fin_seq;
            while (ld_acquire_gpu(&sc->fin_seq) <= fin_seq0) { }
            t_acc[3] += gtime() - tq;
            t_acc[4] += 1ull;
        }
        __syncthreads();
        MD_TRACE(bid == 0 && tid == 0, 6);
    }
    if (bid == 0 && tid == 0 && t_acc[4]) {
        // (the last finalize of this launch is complete: nobody else writes these words)
#pragma unroll
        for (int k = 0; k < 4; ++k) sc->loop_ns[k] += t_acc[k];
        sc->loop_steps += t_acc[4];
    }
}

// flags for the two compacted index lists of the loop: atoms with listed partners, split by "does the list hold a ghost"
__global__ void k_flag_active(int n, const int *__restrict__ nbr_cnt, const int *__restrict__ has_ghost, int want_ghost,
                              int *__restrict__ flag)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int g = has_ghost ? (has_ghost[i] != 0) : 0;
    flag[i] = (nbr_cnt[i] > 0 && g == want_ghost) ? 1 : 0;
}

// idx[base[0] + pos[i]] = i for flagged i; the count of this class goes to count_out[0]
__global__ void k_compact_active(int n, const int *__restrict__ flag, const int *__restrict__ pos, const int *__restrict__ base,
                                 int *__restrict__ idx, int *__restrict__ count_out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = base ? base[0] : 0;
    if (i < n && flag[i]) idx[b + pos[i]] = i;
    if (i == 0) count_out[0] = pos[n];
}

}  // namespace md

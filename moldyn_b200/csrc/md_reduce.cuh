// md_reduce.cuh — K5: deterministic reductions, last-block fold, finalize (T, P, lambda, myu, psi, rebuild decision).
// Part of md_kernels.cuh (included from there, in order; one translation unit).
#pragma once

namespace md {

// ----------------------------------------------------------------------------------------------------
// K5: deterministic reductions.  Lane tree (xor shuffles) → fixed-order sum over warps → one slot per block;
// the last block to finish (atomic ticket) folds the per-block slots in a fixed order and finalizes.
struct Sums {
    double v[NSUM];
};

__device__ __forceinline__ void warp_reduce(Sums &s)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < NSUM - 1; ++q) s.v[q] += __shfl_xor_sync(0xffffffffu, s.v[q], o);
        s.v[NSUM - 1] = fmax(s.v[NSUM - 1], __shfl_xor_sync(0xffffffffu, s.v[NSUM - 1], o));
    }
}

// All threads of the block must call. Result valid in thread 0.
template <int BLOCK>
__device__ __forceinline__ void block_reduce(Sums &s)
{
    __shared__ double sm[BLOCK / 32][NSUM];
    warp_reduce(s);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int q = 0; q < NSUM; ++q) sm[wid][q] = s.v[q];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < BLOCK / 32; ++w) {
#pragma unroll
            for (int q = 0; q < NSUM - 1; ++q) s.v[q] += sm[w][q];
            s.v[NSUM - 1] = fmax(s.v[NSUM - 1], sm[w][NSUM - 1]);
        }
    }
    __syncthreads();
}

// Thermostat coefficient of the NEXT step.  Berendsen (thermostat.rs:31-34): lambda from the temperature at the step
// start.  Nose-Hoover (thermostat.rs:35-39, 59-65): psi advances by half a step with the start temperature, lambda =
// exp(-psi dt/2), then psi advances again with the temperature after the first half-kick (before scaling).
__device__ __forceinline__ double thermostat_lambda(int kind, double dt, double tau, double target, double t_start,
                                                    double t_mid, double &psi)
{
    if (kind == 1) return sqrt(1.0 + dt / tau * (target / t_start - 1.0));
    if (kind == 2) {
        double psi_dot = -((target / t_start) - 1.0) / tau;
        psi += psi_dot * (dt / 2.0);
        const double lambda = exp(-psi * dt / 2.0);
        psi_dot = -((target / t_mid) - 1.0) / tau;
        psi += psi_dot * (dt / 2.0);
        return lambda;
    }
    return 1.0;
}

// Step controls for the first step of a batch from the stored macro state (thermostat.rs:24-44, barostat.rs:21-31)
// plus the displacement bookkeeping that triggers list rebuilds.  psi_in: the caller's Nose-Hoover state.
__device__ __forceinline__ void compute_controls(Scalars *sc, const Params *pr, double psi_in)
{
    double mu = 1.0, psi = psi_in;
    const double lambda = thermostat_lambda(pr->th_kind, pr->dt, pr->th_tau, pr->th_target, sc->temperature,
                                            sc->temperature_mid, psi);
    if (pr->ba_kind == 1) {
        double myu_cubed = 1.0 + pr->dt * pr->ba_beta / pr->ba_tau * (sc->pressure - pr->ba_target);
        mu = cbrt(myu_cubed);
    }
    sc->lambda = lambda;
    sc->mu = mu;
    sc->psi = psi;
    // ΣF = 0, so the COM velocity after the next step's kicks is lambda * vcom: used as the shift that keeps
    // the one-pass thermal sum Σ m|v-c|² free of cancellation.
    sc->shift[0] = sc->vcom[0] * lambda;
    sc->shift[1] = sc->vcom[1] * lambda;
    sc->shift[2] = sc->vcom[2] * lambda;
    // Next drift moves every atom by at most lambda*sqrt(max|v + F c|²)*dt; in build-time units that is
    // multiplied by inv_scale (positions and box have been scaled by Π myu since the build).
    double vmax = lambda * sqrt(sc->max_w2);
    sc->disp_next = vmax * pr->dt * sc->inv_scale;
    // Pair now within r_cut ⇒ at build time within r_cut*inv_scale + 2*disp ≤ r_list must hold.
    double thr = 0.5 * (pr->r_list - pr->r_cut * sc->inv_scale) * (1.0 - 1e-9);
    double d = sc->disp_acc + sc->disp_next;
    sc->need_rebuild = (d > thr) ? 1 : 0;
    if (!(d == d) || !(lambda == lambda) || !(mu == mu) || isinf(d) || isinf(lambda) || isinf(mu)) sc->error = 7;
}

// mode bits of finalize
constexpr int FIN_STEP = 1;  // called at the end of an MD step: commit drift, apply barostat box scaling, count
constexpr int FIN_DIST = 2;  // multi-GPU: publish this rank's sums only; k_finalize_dist finalizes after the all-gather
constexpr int FIN_P2P = 8;   // multi-GPU: exchange the rank sums through the peer mailboxes and finalize right here

// `in` / `pr`: the control words and parameters as they were when the kernel started (the last block copies them into
// shared memory while it waits for the partial sums, so finalize starts without a trip to global memory); `sc`: where the
// results go.  The two may alias (k_finalize_dist).
__device__ __forceinline__ void finalize(Scalars *sc, const Scalars *in, const Params *pr, const Sums &t, int mode)
{
    const double n = (double)pr->n, mass = pr->mass, dt = pr->dt, r_list = pr->r_list, r_cut = pr->r_cut;
    const int th_kind = pr->th_kind, ba_kind = pr->ba_kind;
    const double th_tau = pr->th_tau, th_target = pr->th_target;
    const double ba_beta = pr->ba_beta, ba_tau = pr->ba_tau, ba_target = pr->ba_target;
    double box0 = in->box[0], box1 = in->box[1], box2 = in->box[2];
    const double shift0 = in->shift[0], shift1 = in->shift[1], shift2 = in->shift[2];
    const double lambda_used = in->lambda, mu_used = in->mu;
    double disp_acc = in->disp_acc, inv_scale = in->inv_scale;
    const double disp_next_old = in->disp_next;
    const long long steps_left = in->steps_left, steps_done = in->steps_done;
    double psi = in->psi;

    const double M = n * mass;
    const double vc0 = t.v[0] / M, vc1 = t.v[1] / M, vc2 = t.v[2] / M;  // get_center_of_mass_velocity  mod.rs:12-25
    const double e0 = vc0 - shift0, e1 = vc1 - shift1, e2 = vc2 - shift2;
    const double th2 = t.v[3] - M * (e0 * e0 + e1 * e1 + e2 * e2);       // Σ m |v - vcom|²
    const double thermal = th2 / 2.0;                                    // get_thermal_energy   energy.rs:25-37
    if (mode & FIN_STEP) {
        disp_acc += disp_next_old;  // the drift that preceded this force evaluation
        if (ba_kind == 1) {         // barostat.update: boundary_box *= myu  (barostat.rs:45); x *= myu is deferred
            box0 *= mu_used; box1 *= mu_used; box2 *= mu_used;
            inv_scale /= mu_used;
        }
    }
    const double temperature = (2.0 * thermal) / (3.0 * n * K_B) * 100.0;  // temperature.rs:4-7
    // same for u = v + F c (the state thermostat.update sees after the next first half-kick)
    // (the u sums are only accumulated when something reads them: Nose-Hoover, or a plain force evaluation)
    double temperature_mid = temperature;
    if (th_kind == 2 || !(mode & FIN_STEP)) {
        const double uc0 = t.v[S_MU] / M - shift0, uc1 = t.v[S_MU + 1] / M - shift1, uc2 = t.v[S_MU + 2] / M - shift2;
        const double thu2 = t.v[S_THU] - M * (uc0 * uc0 + uc1 * uc1 + uc2 * uc2);
        temperature_mid = (2.0 * (thu2 / 2.0)) / (3.0 * n * K_B) * 100.0;
    }
    const double volume = box0 * box1 * box2;
    const double pressure = (th2 + (-t.v[5]) * 0.5) / volume / 3.0;         // pressure.rs:5-20
    // controls of the NEXT step (thermostat.rs:24-34, barostat.rs:21-31)
    // (Nose-Hoover's psi is only advanced when this batch has a next step: the first step of the next batch is
    // prepared by k_prepare from the caller's psi.)
    double lambda = 1.0, mu = 1.0;
    const bool more = !(mode & FIN_STEP) || steps_left - 1 > 0;
    if (th_kind == 1 || (th_kind == 2 && more))
        lambda = thermostat_lambda(th_kind, dt, th_tau, th_target, temperature, temperature_mid, psi);
    if (ba_kind == 1) mu = cbrt(1.0 + dt * ba_beta / ba_tau * (pressure - ba_target));
    const double vmax = lambda * sqrt(t.v[S_MAX]);
    const double disp_next = vmax * dt * inv_scale;
    const double thr = 0.5 * (r_list - r_cut * inv_scale) * (1.0 - 1e-9);
    const double d = disp_acc + disp_next;

    for (int k = 0; k < 3; ++k) sc->sum_mv[k] = t.v[k];
    sc->sum_th = t.v[3]; sc->sum_ke = t.v[4]; sc->sum_w = t.v[5]; sc->sum_u = t.v[6]; sc->max_w2 = t.v[S_MAX];
    sc->vcom[0] = vc0; sc->vcom[1] = vc1; sc->vcom[2] = vc2;
    sc->thermal = thermal;
    sc->kinetic = t.v[4] / 2.0;    // get_kinetic_energy   energy.rs:14-22
    sc->potential = t.v[6] / 2.0;  // get_potential_energy energy.rs:40-49
    sc->temperature = temperature;
    sc->temperature_mid = temperature_mid;
    sc->pressure = pressure;
    if (mode & FIN_STEP) sc->psi = psi;
    if (mode & FIN_STEP) {
        sc->lambda_last = lambda_used;
        sc->mu_last = mu_used;
        if (ba_kind == 1) {
            sc->box[0] = box0; sc->box[1] = box1; sc->box[2] = box2;
            sc->mu_pending = mu_used;
        }
        sc->steps_left = steps_left - 1;
        sc->steps_done = steps_done + 1;
        // k_force stored u = v + F*c instead of v unless this was the last step of the batch
        sc->vel_is_half = steps_left - 1 > 0 ? 1 : 0;
        // persistent step loop (md_loop.cuh): every block is past this step's mid-step barrier and face push
        sc->bar_arrive = 0u;
        sc->face_arrive[0] = 0u;
    }
    sc->disp_acc = disp_acc;
    sc->inv_scale = inv_scale;
    sc->lambda = lambda;
    sc->mu = mu;
    // ΣF = 0, so the COM velocity after the next step's kicks is lambda * vcom: the shift that keeps the one-pass
    // thermal sum Σ m|v-c|² free of cancellation.
    sc->shift[0] = vc0 * lambda; sc->shift[1] = vc1 * lambda; sc->shift[2] = vc2 * lambda;
    // Next drift moves every atom by at most lambda*sqrt(max|v + F c|²)*dt (in build-time units: x inv_scale).
    // A pair now within r_cut must have been within r_cut*inv_scale + 2*disp <= r_list at build time.
    sc->disp_next = disp_next;
    sc->need_rebuild = (d > thr) ? 1 : 0;
    if (!(d == d) || !(lambda == lambda) || !(mu == mu) || isinf(d) || isinf(lambda) || isinf(mu)) sc->error = 7;
}

// Last-block epilogue shared by k_force, k_reduce_state and the persistent step loop, in two pieces.
// (1) Every block stores its reduced sums (thread 0 holds them) and takes a ticket; the function returns true in every
//     thread of the block that came last.
template <int BLOCK>
__device__ __forceinline__ bool publish_and_ticket(const Sums &mine, double *__restrict__ partials, Scalars *sc)
{
    __shared__ bool is_last;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NSUM; ++q) __stcg(&partials[(size_t)blockIdx.x * NSUM + q], mine.v[q]);
        // release: the partial sums above and everything this block stored before its last bar.sync; acquire: the block
        // that comes last sees every other block's
        const unsigned int t = atom_add_acq_rel_gpu(&sc->ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    return is_last;
}

// (2) The last block folds the per-block partials in a fixed order, exchanges the rank sums with the other GPUs (FIN_P2P),
//     finalizes the step controls, resets the ticket and release-stores the new sequence number sc->fin_seq — the
//     end-of-step barrier the other blocks of the persistent step loop wait on.
// WARPFOLD (blocks of >= 12 warps, grids of a few hundred blocks — the persistent step loop): warp q sums slot q, lane l
// taking blocks l, l + 32, ... in ascending order, then a lane tree — one L2 round trip and five shuffles; the generic fold's
// serial walk over its BLOCK/6 groups costs 2.4 us at 512 threads (measured, profiles/r02_loop_trace_v1.txt).
template <int BLOCK, bool WARPFOLD = false>
__device__ __forceinline__ void last_block_finalize(double *__restrict__ partials, Scalars *sc, const Params *pr, int mode,
                                                    const Peers *peers_p)
{
    static_assert(!WARPFOLD || BLOCK >= 32 * NSUM, "one warp per slot");
    MD_TRACE(threadIdx.x == 0, 8);
    // (a) A copy of the control words and parameters finalize reads goes to shared memory — those loads are in flight
    // together with (b) the fold of the per-block partials: thread (g, q) adds slot q of blocks g, g+G, g+2G, … in ascending
    // order (independent loads, one L2 round trip), then the G group sums of a slot are added in group order.
    // Fixed assignment, fixed order: the result depends on the grid size only.
    constexpr int H = NSUM / 2;   // slot pairs: 128-bit loads
    constexpr int G = BLOCK / H;  // groups of blocks
    static_assert(NSUM % 2 == 0, "slots are folded in pairs");
    __shared__ double fold[WARPFOLD ? 1 : G][NSUM];
    __shared__ double folded[NSUM];
    __shared__ Scalars sc_in;
    __shared__ Params pr_in;
    {
        constexpr int WS = (int)(sizeof(Scalars) / 8), WP = (int)(sizeof(Params) / 8);
        static_assert(sizeof(Scalars) % 8 == 0 && sizeof(Params) % 8 == 0, "copied as 64-bit words");
        const unsigned long long *gs = reinterpret_cast<const unsigned long long *>(sc);
        const unsigned long long *gp = reinterpret_cast<const unsigned long long *>(pr);
        unsigned long long *ss_ = reinterpret_cast<unsigned long long *>(&sc_in), *sp_ = reinterpret_cast<unsigned long long *>(&pr_in);
        for (int w = threadIdx.x; w < WS + WP; w += BLOCK) {
            if (w < WS) ss_[w] = __ldcg(gs + w);
            else sp_[w - WS] = __ldcg(gp + (w - WS));
        }
    }
    if (WARPFOLD) {
        const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
        if (q < NSUM) {
            double v = 0.0;
            for (unsigned b = lane; b < gridDim.x; b += 32) {
                const double w = __ldcg(partials + (size_t)b * NSUM + q);
                v = q == NSUM - 1 ? fmax(v, w) : v + w;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double w = __shfl_xor_sync(0xffffffffu, v, o);
                v = q == NSUM - 1 ? fmax(v, w) : v + w;
            }
            if (lane == 0) folded[q] = v;
        }
    } else {
        const int h = threadIdx.x % H, g = threadIdx.x / H;
        if (g < G) {
            const bool has_max = (h == H - 1);  // the last slot of the last pair is the running maximum
            double ax = 0.0, ay = 0.0;
            const double2 *src = reinterpret_cast<const double2 *>(partials) + h;
            constexpr int U = 16;              // loads in flight per thread
            unsigned int b = g;
            for (; b + (U - 1) * G < gridDim.x; b += U * G) {
                double2 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) v[u] = __ldcg(src + (size_t)(b + u * G) * H);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    ax += v[u].x;
                    ay = has_max ? fmax(ay, v[u].y) : ay + v[u].y;
                }
            }
            {
                double2 v[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned int bb = b + u * G;
                    v[u] = bb < gridDim.x ? __ldcg(src + (size_t)bb * H) : make_double2(0.0, 0.0);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    ax += v[u].x;
                    ay = has_max ? fmax(ay, v[u].y) : ay + v[u].y;
                }
            }
            fold[g][2 * h] = ax;
            fold[g][2 * h + 1] = ay;
        }
    }
    __syncthreads();
    if (!WARPFOLD && threadIdx.x < NSUM) {
        const int q = threadIdx.x;
        double a = fold[0][q];
        for (int g = 1; g < G; ++g) a = (q == NSUM - 1) ? fmax(a, fold[g][q]) : a + fold[g][q];
        folded[q] = a;
    }
    if (!WARPFOLD) __syncthreads();
    Sums acc;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < NSUM; ++q) acc.v[q] = folded[q];
    }
    MD_TRACE(threadIdx.x == 0, 9);
    const unsigned long long t_last = gtime();  // every block has finished its atoms
    if (mode & FIN_P2P) {
        // All-gather of the rank sums through peer memory, fused into this kernel.  Every rank stores its 12 sums into every
        // rank's mailbox as 24 self-flagged 8-byte words (Mail::ll) — no fence, no separate flag: one NVLink store latency —
        // polls its own mailbox until every rank's words carry this reduction's sequence number, and folds the ranks in rank
        // order: identical lambda / myu / rebuild decision everywhere.
        constexpr int W = 2 * NSUM;
        __shared__ double my_sums[NSUM];
        __shared__ unsigned int parts[MAX_PEERS][W];
        __shared__ int timed_out;
        const Peers &peers = *peers_p;
        const unsigned long long seq = sc_in.epoch + 1;
        const int buf = (int)(seq & 1ull);
        const unsigned long long tag = (seq & 0xffffffffull) << 32;
        if (threadIdx.x == 0) {
#pragma unroll
            for (int q = 0; q < NSUM; ++q) my_sums[q] = acc.v[q];
            timed_out = 0;
        }
        __syncthreads();
        const unsigned long long t_wait = gtime();
        for (int t = threadIdx.x; t < peers.nranks * W; t += BLOCK) {
            const int r = t / W, w = t - r * W;
            const unsigned long long bits = (unsigned long long)__double_as_longlong(my_sums[w >> 1]);
            const unsigned long long part = (w & 1) ? (bits >> 32) : (bits & 0xffffffffull);
            st_relaxed_sys(&peers.mail[r]->ll[buf][peers.rank][w], tag | part);
        }
        const Mail *own = peers.mail[peers.rank];
        for (int t = threadIdx.x; t < peers.nranks * W; t += BLOCK) {
            const int r = t / W, w = t - r * W;
            unsigned long long v = ld_relaxed_sys(&own->ll[buf][r][w]);
            if ((v & 0xffffffff00000000ull) != tag) {
                const unsigned long long t0 = gtime();
                for (;;) {
                    v = ld_relaxed_sys(&own->ll[buf][r][w]);
                    if ((v & 0xffffffff00000000ull) == tag) break;
                    if (gtime() - t0 > 20000000000ull) { timed_out = 1; break; }  // a peer died or the ranks diverged
                }
            }
            parts[r][w] = (unsigned int)(v & 0xffffffffull);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            sc->wait_sums_ns = sc_in.wait_sums_ns + (gtime() - t_wait);
            Sums t;
#pragma unroll
            for (int q = 0; q < NSUM; ++q) t.v[q] = 0.0;
            for (int r = 0; r < peers.nranks; ++r) {
#pragma unroll
                for (int q = 0; q < NSUM; ++q) {
                    const double v = __longlong_as_double((long long)(((unsigned long long)parts[r][2 * q + 1] << 32) | parts[r][2 * q]));
                    t.v[q] = q == NSUM - 1 ? fmax(t.v[q], v) : t.v[q] + v;
                }
            }
            finalize(sc, &sc_in, &pr_in, t, mode);
            if (timed_out) sc->error = 3;  // MD_ERR_NCCL: a peer never delivered
            if (sc_in.t_start != ~0ull) sc->force_atoms_ns = sc_in.force_atoms_ns + (t_last - sc_in.t_start);
            sc->force_tail_ns = sc_in.force_tail_ns + (gtime() - t_last);
            sc->t_start = ~0ull;
            sc->epoch = seq;
            sc->ticket = 0;
            st_release_gpu(&sc->fin_seq, sc_in.fin_seq + 1);
        }
        return;
    }
    if (threadIdx.x == 0) {
        if (mode & FIN_DIST) {
#pragma unroll
            for (int q = 0; q < NSUM; ++q) sc->rank_sums[q] = acc.v[q];
            sc->ticket = 0;
            return;
        }
        finalize(sc, &sc_in, &pr_in, acc, mode);
        sc->ticket = 0;
        MD_TRACE(true, 10);
        // everything above is visible to whoever acquires the new sequence number
        st_release_gpu(&sc->fin_seq, sc_in.fin_seq + 1);
        MD_TRACE(true, 11);
    }
}

template <int BLOCK>
__device__ __forceinline__ void grid_reduce_finalize(Sums &mine, double *__restrict__ partials, Scalars *sc,
                                                     const Params *pr, int mode, const Peers *peers_p)
{
    if (!publish_and_ticket<BLOCK>(mine, partials, sc)) return;
    last_block_finalize<BLOCK>(partials, sc, pr, mode, peers_p);
}

// Adds one atom's terms. (wx,wy,wz) = v + F*c is the velocity the next kick_drift moves this atom with (before lambda).
__device__ __forceinline__ void accumulate_sums(Sums &s, double m, double vx, double vy, double vz, double wx,
                                                double wy, double wz, double w, double u, const double *shift)
{
    s.v[0] += m * vx; s.v[1] += m * vy; s.v[2] += m * vz;
    double ax = vx - shift[0], ay = vy - shift[1], az = vz - shift[2];
    s.v[S_TH] += m * (ax * ax + ay * ay + az * az);
    s.v[S_KE] += m * (vx * vx + vy * vy + vz * vz);
    s.v[S_W] += w;
    s.v[S_U] += u;
    s.v[S_MU] += m * wx; s.v[S_MU + 1] += m * wy; s.v[S_MU + 2] += m * wz;
    double bx = wx - shift[0], by = wy - shift[1], bz = wz - shift[2];
    s.v[S_THU] += m * (bx * bx + by * by + bz * bz);
    s.v[S_MAX] = fmax(s.v[S_MAX], wx * wx + wy * wy + wz * wz);
}

// Standalone K5 over the stored state (after upload, or when only the macro parameters are wanted).
constexpr int RED_BLOCK = 256;
__global__ void __launch_bounds__(RED_BLOCK) k_reduce_state(int n, Arrays a, double *__restrict__ partials,
                                                            Scalars *sc, const Params *__restrict__ pr, int mode)
{
    Sums s;
#pragma unroll
    for (int q = 0; q < NSUM; ++q) s.v[q] = 0.0;
    const double shift[3] = {sc->shift[0], sc->shift[1], sc->shift[2]};
    const double c = pr->half_dt_m, m = pr->mass;
    for (int i = blockIdx.x * RED_BLOCK + threadIdx.x; i < n; i += gridDim.x * RED_BLOCK) {
        double vx = a.vx[i], vy = a.vy[i], vz = a.vz[i];
        double wx = __dadd_rn(vx, __dmul_rn(a.fx[i], c)), wy = __dadd_rn(vy, __dmul_rn(a.fy[i], c)),
               wz = __dadd_rn(vz, __dmul_rn(a.fz[i], c));
        accumulate_sums(s, m, vx, vy, vz, wx, wy, wz, a.w[i], a.u[i], shift);
    }
    block_reduce<RED_BLOCK>(s);
    grid_reduce_finalize<RED_BLOCK>(s, partials, sc, pr, mode, nullptr);
}

}  // namespace md

// md_loop.cuh — the persistent step loop of dilute systems: ONE cooperative kernel runs MD steps until the neighbour list
// has to be rebuilt (or the batch ends), with two grid-wide synchronisations per step instead of two kernel boundaries and
// with the velocities resident in shared memory for the whole launch.
// Part of md_kernels.cuh (included from there, in order; one translation unit).
#pragma once

namespace md {

// Why.  Round-1 measurements: the two-kernel step moves 152 B per atom and step through L2/HBM (x and u
// are read by both kernels) and, below ~10^5 atoms per GPU (the 8-GPU slabs, C1, C2), is nothing but fixed latencies — two
// launches, a 592-slot fold, a mailbox exchange.  A thread of this kernel owns the same pairs of atoms in every phase of every
// step of the launch, so what only the owner ever touches — the velocity u — never leaves the SM: 24 B/atom in shared memory
// (10^6 atoms: 162 KB of the 227 KB per SM), loaded once per launch and written back when the loop exits.  A step
// (integrator.rs:14-59) is then
//
//   phase A  lambda-scale, pending barostat scale, drift, wrap of the thread's atoms (k_kick_drift's arithmetic): read x,
//            write x'.  Atoms WITHOUT listed partners (60-80 % of the gas) finish their step here: F = 0, v'' = u' =
//            lambda*u stays in shared memory, their K5 terms are summed.
//   barrier  every drifted position is visible to the whole grid.  Multi-GPU: block 0 then raises the step's flag in both
//            neighbours' mailboxes (one release store each over NVLink) — nothing else crosses the link on the sender's side.
//   phase B  the thread's pairs with listed partners: pair forces (gathers from L2), both half-kicks, K5 terms; u' back to
//            shared memory.  Multi-GPU: pairs whose lists hold no ghost first; the others wait for the neighbours' flags
//            only when the block gets there and then read their ghost partners STRAIGHT FROM THE NEIGHBOUR'S PLANES (peer
//            loads): a ghost is a face atom of the neighbour — a prefix / suffix of its sorted order — so its address is
//            the list index plus a constant.  (A first version pushed the face atoms into the neighbours' ghost slots from
//            phase A: the pushing blocks' system-scope fences sat on the critical path, drift + barrier 17.9 us at
//            5*10^5 atoms per GPU against 12.2 us for 10^6 on one GPU — profiles/r02_bench_c3_2gpu_push.txt.)
//   tail     block sums -> ticket -> the last block folds, exchanges the rank sums through the peer mailboxes, finalizes
//            (T, P, lambda, myu, rebuild decision) and release-stores the step's sequence number; the other blocks wait for
//            it and start the next step.
//
// Per step and atom the kernel reads x (24 B), writes x' (24 B) and re-reads x' plus the list head (32 B, L2 hits) —
// against 152 B of the two-kernel step.  Per-atom arithmetic is the same drift_one() / pair term / half-kick sequence.
// Everything another block (or GPU) wrote during the launch is read with ld.global.cg (L2): the L1 is not coherent across SMs.
//
// (A first version walked a compacted list of the atoms with partners, one atom per lane, in phase B.  Measured on B200
// (profiles/r02_loop_trace_v1.txt): 20.7 us for 3.5e5 active atoms — every scattered 8-byte access pulls a 32-byte sector,
// the phase moved 4x the bytes of a coalesced pass.  Same finding as round 1's k_force_sparse.)
constexpr int LOOP_BLOCK = 512;
constexpr int LOOP_MAX_PAIRS = 9;         // pairs of atoms per thread whose velocities fit in shared memory (9 x 24 KB per block)
constexpr int LOOP_GHOST_FLAG = 1 << 30;  // in the loop's copy of the list counts: the list holds a ghost atom

struct LoopArgs {
    int n;                  // owned atoms
    int npad, cap;          // row stride and capacity of the neighbour table
    int pairs_per_thread;   // P: thread (b, l) owns pairs (b*P + p)*LOOP_BLOCK + l, p < P
    int pairs_in_smem;      // PS <= P: the velocities of pairs p < PS live in shared memory, the others in the planes
    Arrays a;
    const int *nbr;
    const int *cntg;        // list counts (| LOOP_GHOST_FLAG on the multi-GPU path)
    const int *bnd_pairs;   // multi-GPU: the pairs with a ghost partner, compacted at the last rebuild, and (device) their number:
    const int *n_bnd;       // they are spread over the whole grid in the second pass of the force phase (their velocities stay
                            // in the planes) — each costs a dependent chain of NVLink reads, and the slab's faces would
                            // otherwise pile them all onto the first and the last blocks
    double *partials;
    Scalars *sc;
    const Params *pr;
    const Peers *peers;     // NULL on one GPU
    // multi-GPU: ghost j in [n, n_gl) is atom j + off of the LEFT neighbour's planes gl*, ghost j >= n_gl of the right one's
    int n_gl;
    const double *glx, *gly, *glz, *grx, *gry, *grz;  // (already offset: index them with the list entry j)
    long long max_steps;    // steps this launch may run (host-stepped loop: 1)
    ForceConsts fc;
};

struct LoopCtl {
    double lambda, mup, Lx, Ly, Lz, shift[3];
    long long steps_left;
    unsigned long long fin_seq, epoch;
    int halted, half;
};

// finish_atom() with the running sums in registers (same operations, same order per atom)
__device__ __forceinline__ void finish_atom_r(Sums &s, const PairAcc &f, double &vx, double &vy, double &vz, double lambda,
                                              double c, double mass, const double *shift, double &wx, double &wy, double &wz,
                                              bool nh)
{
    vx = __dadd_rn(__dmul_rn(vx, lambda), __dmul_rn(f.fx, c));  // v'' = lambda*u + F*c
    vy = __dadd_rn(__dmul_rn(vy, lambda), __dmul_rn(f.fy, c));
    vz = __dadd_rn(__dmul_rn(vz, lambda), __dmul_rn(f.fz, c));
    wx = __dadd_rn(vx, __dmul_rn(f.fx, c));                     // u' = v'' + F*c
    wy = __dadd_rn(vy, __dmul_rn(f.fy, c));
    wz = __dadd_rn(vz, __dmul_rn(f.fz, c));
    s.v[0] += mass * vx; s.v[1] += mass * vy; s.v[2] += mass * vz;
    const double ax = vx - shift[0], ay = vy - shift[1], az = vz - shift[2];
    s.v[S_TH] += mass * (ax * ax + ay * ay + az * az);
    s.v[S_KE] += mass * (vx * vx + vy * vy + vz * vz);
    s.v[S_W] += f.w;
    s.v[S_U] += f.u;
    if (nh) {
        s.v[S_MU] += mass * wx; s.v[S_MU + 1] += mass * wy; s.v[S_MU + 2] += mass * wz;
        const double bx = wx - shift[0], by = wy - shift[1], bz = wz - shift[2];
        s.v[S_THU] += mass * (bx * bx + by * by + bz * bz);
    }
    s.v[S_MAX] = fmax(s.v[S_MAX], wx * wx + wy * wy + wz * wz);
}

// Block sums for a 16-warp block: lane tree, then the first warp folds the 16 warp results with a second lane tree (the
// serial fold of block_reduce<> costs 1.9 us at 512 threads — measured, profiles/r02_loop_trace_v1.txt).  Result in thread 0.
// The trees are shuffle-bound (12 sums x 5 levels x 2 words per warp, 16 warps per SM), so only the sums something reads this
// step go through them (`mask`, grid-uniform): momentum, thermal energy and the largest speed always; the virial when a
// barostat runs; kinetic and potential energy only on a step that leaves the State behind; the u sums for Nose-Hoover.
__device__ __forceinline__ void block_reduce_loop(Sums &s, unsigned int mask)
{
    static_assert(LOOP_BLOCK == 512, "16 warps");
    __shared__ double sm[NSUM][16];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NSUM; ++q) {
        if (!((mask >> q) & 1u)) continue;
        double v = s.v[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double w = __shfl_xor_sync(0xffffffffu, v, o);
            v = q == NSUM - 1 ? fmax(v, w) : v + w;
        }
        if (lane == 0) sm[q][wid] = v;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int q = 0; q < NSUM; ++q) {
            if (!((mask >> q) & 1u)) continue;
            double v = lane < 16 ? sm[q][lane] : 0.0;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const double w = __shfl_xor_sync(0xffffffffu, v, o);
                v = q == NSUM - 1 ? fmax(v, w) : v + w;
            }
            s.v[q] = v;
        }
    }
    __syncthreads();
}

template <bool EXACT, bool USMEM>
__global__ void __launch_bounds__(LOOP_BLOCK, 1) k_md_loop(const LoopArgs A)
{
    extern __shared__ __align__(16) unsigned char loop_smem[];
    double2 *su = reinterpret_cast<double2 *>(loop_smem);  // [3][P][LOOP_BLOCK]: ux, uy, uz of the thread's pairs
    __shared__ LoopCtl ctl;
    __shared__ unsigned long long t_acc[5];  // block 0's phase clocks + step count (thread 0)
    const int tid = threadIdx.x, bid = blockIdx.x, nb = gridDim.x;
    Scalars *sc = A.sc;
    const Params *__restrict__ pr = A.pr;
    const Arrays &a = A.a;
    const ForceConsts &fc = A.fc;
    const int n = A.n, P = A.pairs_per_thread, PS = USMEM ? A.pairs_in_smem : 0;
    const int npairs = (n + 1) >> 1;
    const bool multi = A.peers != nullptr;
    const bool nh = pr->th_kind == 2;
    const double dt = pr->dt, hc = fc.hc, mass = fc.mass;
    const PairAcc zero = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (tid == 0) { t_acc[0] = t_acc[1] = t_acc[2] = t_acc[3] = t_acc[4] = 0ull; }
    // pair p of this thread, and where its velocity lives
    auto pair_of = [&](int p) { return (bid * P + p) * LOOP_BLOCK + tid; };
    auto ld_u = [&](bool in_smem, int p, int t, double2 &ux, double2 &uy, double2 &uz) {
        if (USMEM && in_smem) {
            ux = su[(0 * PS + p) * LOOP_BLOCK + tid]; uy = su[(1 * PS + p) * LOOP_BLOCK + tid]; uz = su[(2 * PS + p) * LOOP_BLOCK + tid];
        } else {
            ux = __ldcg(reinterpret_cast<const double2 *>(a.vx) + t); uy = __ldcg(reinterpret_cast<const double2 *>(a.vy) + t);
            uz = __ldcg(reinterpret_cast<const double2 *>(a.vz) + t);
        }
    };
    auto st_u = [&](bool in_smem, int p, int t, bool has1, const double2 &ux, const double2 &uy, const double2 &uz) {
        if (USMEM && in_smem) {
            su[(0 * PS + p) * LOOP_BLOCK + tid] = ux; su[(1 * PS + p) * LOOP_BLOCK + tid] = uy; su[(2 * PS + p) * LOOP_BLOCK + tid] = uz;
        } else if (has1) {
            reinterpret_cast<double2 *>(a.vx)[t] = ux; reinterpret_cast<double2 *>(a.vy)[t] = uy;
            reinterpret_cast<double2 *>(a.vz)[t] = uz;
        } else {  // odd tail: the slot after it may belong to a ghost atom
            a.vx[2 * t] = ux.x; a.vy[2 * t] = uy.x; a.vz[2 * t] = uz.x;
        }
    };
    bool loaded = false;

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
        const bool store_state = ctl.steps_left <= 1;  // last step of the batch: the resident State must be complete
        if (!loaded) {
            // first step of the launch: velocities into shared memory, with the first half-kick of a batch
            // (integrator.rs:28-34) when the planes hold v and not u = v + F c
            const bool half = ctl.half != 0;
            for (int p = 0; p < P; ++p) {
                const int t = pair_of(p);
                if (t >= npairs) break;
                const bool has1 = 2 * t + 1 < n;
                const int2 C = reinterpret_cast<const int2 *>(A.cntg)[t];
                const bool in_smem = p < PS && (((C.x | (has1 ? C.y : 0)) & LOOP_GHOST_FLAG) == 0);
                double2 ux = __ldcg(reinterpret_cast<const double2 *>(a.vx) + t), uy = __ldcg(reinterpret_cast<const double2 *>(a.vy) + t),
                        uz = __ldcg(reinterpret_cast<const double2 *>(a.vz) + t);
                if (!half) {
                    const double2 fx = __ldcg(reinterpret_cast<const double2 *>(a.fx) + t),
                                  fy = __ldcg(reinterpret_cast<const double2 *>(a.fy) + t),
                                  fz = __ldcg(reinterpret_cast<const double2 *>(a.fz) + t);
                    ux.x = __dadd_rn(ux.x, __dmul_rn(fx.x, hc)); ux.y = __dadd_rn(ux.y, __dmul_rn(fx.y, hc));
                    uy.x = __dadd_rn(uy.x, __dmul_rn(fy.x, hc)); uy.y = __dadd_rn(uy.y, __dmul_rn(fy.y, hc));
                    uz.x = __dadd_rn(uz.x, __dmul_rn(fz.x, hc)); uz.y = __dadd_rn(uz.y, __dmul_rn(fz.y, hc));
                }
                if ((USMEM && in_smem) || !half) st_u(in_smem, p, t, has1, ux, uy, uz);
            }
            loaded = true;
        }
        Sums s;
#pragma unroll
        for (int q = 0; q < NSUM; ++q) s.v[q] = 0.0;

        // ---- phase A ----------------------------------------------------------------------------------------------------
        auto drift_pair = [&](int p, int t) {
            const int i0 = 2 * t;
            const bool has1 = i0 + 1 < n;
            double2 x = __ldcg(reinterpret_cast<const double2 *>(a.x) + t), y = __ldcg(reinterpret_cast<const double2 *>(a.y) + t),
                    z = __ldcg(reinterpret_cast<const double2 *>(a.z) + t);
            int2 C = reinterpret_cast<const int2 *>(A.cntg)[t];
            if (!has1) C.y = 1;  // (not an atom: nothing to finish, no ghost flag)
            const bool in_smem = p < PS && ((C.x | C.y) & LOOP_GHOST_FLAG) == 0;
            double2 ux, uy, uz;
            ld_u(in_smem, p, t, ux, uy, uz);
            drift_one(x.x, ux.x, lambda, mup, dt, Lx); drift_one(x.y, ux.y, lambda, mup, dt, Lx);
            drift_one(y.x, uy.x, lambda, mup, dt, Ly); drift_one(y.y, uy.y, lambda, mup, dt, Ly);
            drift_one(z.x, uz.x, lambda, mup, dt, Lz); drift_one(z.y, uz.y, lambda, mup, dt, Lz);
            if (has1) {
                reinterpret_cast<double2 *>(a.x)[t] = x; reinterpret_cast<double2 *>(a.y)[t] = y;
                reinterpret_cast<double2 *>(a.z)[t] = z;
            } else {
                a.x[i0] = x.x; a.y[i0] = y.x; a.z[i0] = z.x;
            }
            // atoms without listed partners: F = 0, the step ends here (v'' = u' = lambda*u); the others keep u for phase B
            const bool s0 = C.x == 0, s1 = C.y == 0;
            if (s0) {
                double wx, wy, wz;
                finish_atom_r(s, zero, ux.x, uy.x, uz.x, lambda, hc, mass, shift, wx, wy, wz, nh);
                ux.x = wx; uy.x = wy; uz.x = wz;
            }
            if (s1) {
                double wx, wy, wz;
                finish_atom_r(s, zero, ux.y, uy.y, uz.y, lambda, hc, mass, shift, wx, wy, wz, nh);
                ux.y = wx; uy.y = wy; uz.y = wz;
            }
            if (s0 || s1) st_u(in_smem, p, t, has1, ux, uy, uz);
            if (store_state) {
                if (s0 && s1) {
                    const double2 z2 = make_double2(0.0, 0.0);
                    reinterpret_cast<double2 *>(a.fx)[t] = z2; reinterpret_cast<double2 *>(a.fy)[t] = z2;
                    reinterpret_cast<double2 *>(a.fz)[t] = z2; reinterpret_cast<double2 *>(a.u)[t] = z2;
                    reinterpret_cast<double2 *>(a.w)[t] = z2;
                } else {
                    if (s0) { a.fx[i0] = 0.0; a.fy[i0] = 0.0; a.fz[i0] = 0.0; a.u[i0] = 0.0; a.w[i0] = 0.0; }
                    if (s1 && has1) { a.fx[i0 + 1] = 0.0; a.fy[i0 + 1] = 0.0; a.fz[i0 + 1] = 0.0; a.u[i0 + 1] = 0.0; a.w[i0 + 1] = 0.0; }
                }
            }
        };
        for (int p = 0; p < P; ++p) {
            const int t = pair_of(p);
            if (t < npairs) drift_pair(p, t);
        }
        if (tid == 0) { const unsigned long long t = gtime(); t_acc[0] += t - tq; tq = t; }
        MD_TRACE(bid == 0 && tid == 0, 1);

        // ---- mid-step barrier: all drifted positions are visible grid-wide -------------------------------------------------
        __syncthreads();
        if (nb > 1 && tid == 0) {
            // release: this block's drifted positions (the bar.sync above makes the release cumulative over its threads);
            // the acquiring loads also invalidate this SM's L1
            atom_add_release_gpu(&sc->bar_arrive, 1u);
            while (ld_acquire_gpu(&sc->bar_arrive) < (unsigned int)nb) { }
        }
        __syncthreads();
        if (multi && bid == nb - 1 && tid == 0) {
            // every block's drifted positions are in this GPU's L2 (the barrier's releases): the neighbours may read our face
            // atoms of this step.  (The last block has the least work: the system-scope release stalls its first warp for ~3 us.)
            // ONE system-scope fence, then the two flags as relaxed stores (a release store each would pay the fence twice:
            // 7.4 us of halo wait on the neighbour, profiles/r02_bench_c3_2gpu_pull_v1.txt)
            __threadfence_system();
            st_relaxed_sys(&A.peers->mail[A.peers->left]->halo_seq[1], ctl.epoch + 1);   // we are its right side
            st_relaxed_sys(&A.peers->mail[A.peers->right]->halo_seq[0], ctl.epoch + 1);
        }
        if (tid == 0) { const unsigned long long t = gtime(); t_acc[1] += t - tq; tq = t; }
        MD_TRACE(bid == 0 && tid == 0, 2);

        // ---- phase B ----------------------------------------------------------------------------------------------------
        LjConst c;
        c.Lx = Lx; c.Ly = Ly; c.Lz = Lz;
        c.hx = Lx / 2.0; c.hy = Ly / 2.0; c.hz = Lz / 2.0;
        c.hxi = __double2hiint(c.hx); c.hyi = __double2hiint(c.hy); c.hzi = __double2hiint(c.hz);
        const size_t stride = (size_t)A.npad;
        auto force_pair = [&](int p, int t, int2 C, auto ghosts_tag) {
            constexpr bool ghosts = decltype(ghosts_tag)::value;  // compile-time: the owned-partner pass keeps plain base pointers
            const bool in_smem = !ghosts && p < PS;
            const int i0 = 2 * t;
            const bool has1 = i0 + 1 < n;
            // (positions are stable throughout the phase and the barrier's acquiring loads invalidated this SM's L1 — CCTL.IVALL
            // in the SASS — so the gathers may use it: partners shared by neighbouring atoms hit)
            const double2 X = reinterpret_cast<const double2 *>(a.x)[t], Y = reinterpret_cast<const double2 *>(a.y)[t],
                          Z = reinterpret_cast<const double2 *>(a.z)[t];
            int2 J = reinterpret_cast<const int2 *>(A.nbr)[t];  // row 0
            double2 ux, uy, uz;
            ld_u(in_smem, p, t, ux, uy, uz);
            PairAcc f0 = zero, f1 = zero;
            const int kmax = max(C.x, C.y);
            if constexpr (!ghosts) {
                // owned partners: one list row per trip, the next row's indices fetched while the current gathers fly
                for (int k = 0; k < kmax; ++k) {
                    const bool a0 = k < C.x, a1 = k < C.y;
                    const int j0 = a0 ? J.x : i0, j1 = a1 ? J.y : i0;
                    if (k + 1 < kmax) J = *reinterpret_cast<const int2 *>(A.nbr + (size_t)(k + 1) * stride + i0);
                    const double xa = a.x[j0], ya = a.y[j0], za = a.z[j0];
                    const double xb = a.x[j1], yb = a.y[j1], zb = a.z[j1];
                    if (EXACT) {
                        if (a0) pair_exact(f0, xa, ya, za, X.x, Y.x, Z.x, c, fc);
                        if (a1) pair_exact(f1, xb, yb, zb, X.y, Y.y, Z.y, c, fc);
                    } else {
                        pair_fast_branchy(f0, a0, xa, ya, za, X.x, Y.x, Z.x, c, fc);
                        pair_fast_branchy(f1, a1, xb, yb, zb, X.y, Y.y, Z.y, c, fc);
                    }
                }
            } else {
                // pairs with ghost partners: TWO list rows per trip — both rows' indices, then the four partners' coordinates,
                // then the pair terms: a dependent round trip here is an NVLink read (~2.5 us), and most of these atoms have
                // at most two partners.  A partner at or beyond n lives on a neighbour and is read there.
                int2 Jb = kmax > 1 ? *reinterpret_cast<const int2 *>(A.nbr + stride + i0) : J;
                for (int k = 0; k < kmax; k += 2) {
                    const bool a0 = k < C.x, a1 = k < C.y, b0 = k + 1 < C.x, b1 = k + 1 < C.y;
                    const int jj[4] = {a0 ? J.x : i0, a1 ? J.y : i0, b0 ? Jb.x : i0, b1 ? Jb.y : i0};
                    if (k + 2 < kmax) J = *reinterpret_cast<const int2 *>(A.nbr + (size_t)(k + 2) * stride + i0);
                    if (k + 3 < kmax) Jb = *reinterpret_cast<const int2 *>(A.nbr + (size_t)(k + 3) * stride + i0);
                    double xq[4], yq[4], zq[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int j = jj[u];
                        const bool own = j < n, l = j < A.n_gl;
                        xq[u] = (own ? a.x : (l ? A.glx : A.grx))[j];
                        yq[u] = (own ? a.y : (l ? A.gly : A.gry))[j];
                        zq[u] = (own ? a.z : (l ? A.glz : A.grz))[j];
                    }
                    if (EXACT) {
                        if (a0) pair_exact(f0, xq[0], yq[0], zq[0], X.x, Y.x, Z.x, c, fc);
                        if (a1) pair_exact(f1, xq[1], yq[1], zq[1], X.y, Y.y, Z.y, c, fc);
                        if (b0) pair_exact(f0, xq[2], yq[2], zq[2], X.x, Y.x, Z.x, c, fc);
                        if (b1) pair_exact(f1, xq[3], yq[3], zq[3], X.y, Y.y, Z.y, c, fc);
                    } else {
                        pair_fast_branchy(f0, a0, xq[0], yq[0], zq[0], X.x, Y.x, Z.x, c, fc);
                        pair_fast_branchy(f1, a1, xq[1], yq[1], zq[1], X.y, Y.y, Z.y, c, fc);
                        pair_fast_branchy(f0, b0, xq[2], yq[2], zq[2], X.x, Y.x, Z.x, c, fc);
                        pair_fast_branchy(f1, b1, xq[3], yq[3], zq[3], X.y, Y.y, Z.y, c, fc);
                    }
                }
            }
            double2 wx = ux, wy = uy, wz = uz;
            if (C.x > 0) finish_atom_r(s, f0, ux.x, uy.x, uz.x, lambda, hc, mass, shift, wx.x, wy.x, wz.x, nh);
            if (C.y > 0) finish_atom_r(s, f1, ux.y, uy.y, uz.y, lambda, hc, mass, shift, wx.y, wy.y, wz.y, nh);
            // (an atom of the pair without partners was finished in phase A: ux == wx holds its lambda*u already)
            if (store_state) {
                if (C.x > 0) { a.fx[i0] = f0.fx; a.fy[i0] = f0.fy; a.fz[i0] = f0.fz; a.u[i0] = f0.u; a.w[i0] = f0.w; }
                if (C.y > 0) { a.fx[i0 + 1] = f1.fx; a.fy[i0 + 1] = f1.fy; a.fz[i0 + 1] = f1.fz; a.u[i0 + 1] = f1.u; a.w[i0 + 1] = f1.w; }
                st_u(in_smem, p, t, has1, ux, uy, uz);   // v''
            } else {
                st_u(in_smem, p, t, has1, wx, wy, wz);   // u'
            }
        };
        for (int p = 0; p < P; ++p) {  // the thread's own pairs whose partners are all owned
            const int t = pair_of(p);
            if (t >= npairs) break;
            int2 C = reinterpret_cast<const int2 *>(A.cntg)[t];
            if (2 * t + 1 >= n) C.y = 0;
            if (((C.x | C.y) & LOOP_GHOST_FLAG) == 0 && (C.x | C.y) != 0) force_pair(p, t, C, std::false_type{});
        }
        const int n_bnd = multi ? A.n_bnd[0] : 0;
        if (bid * LOOP_BLOCK < n_bnd) {
            // Only blocks with pairs to do here wait for the neighbours' flags: their drifted positions of this step are in
            // their L2 (the acquiring loads also drop this SM's L1 lines of the previous step's peer reads).  A rank without
            // such pairs does not wait at all; what keeps a neighbour from overwriting positions this rank is still reading
            // is the sum exchange of the tail — no rank leaves step k before every rank has finished its force phase.
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
            for (int k = bid * LOOP_BLOCK + tid; k < n_bnd; k += nb * LOOP_BLOCK) {
                const int t = A.bnd_pairs[k];
                int2 C = reinterpret_cast<const int2 *>(A.cntg)[t];
                if (2 * t + 1 >= n) C.y = 0;
                C.x &= ~LOOP_GHOST_FLAG; C.y &= ~LOOP_GHOST_FLAG;
                force_pair(0, t, C, std::true_type{});
            }
        }
        if (tid == 0) { const unsigned long long t = gtime(); t_acc[2] += t - tq; tq = t; }
        MD_TRACE(bid == 0 && tid == 0, 3);

        // ---- tail: block sums, ticket, last block folds + finalizes + releases -------------------------------------------
        {
            unsigned int mask = 0xfu | (1u << S_MAX);                        // S_MV x3, S_TH, S_MAX
            if (store_state) mask |= (1u << S_KE) | (1u << S_U) | (1u << S_W);
            if (pr->ba_kind == 1) mask |= 1u << S_W;
            if (nh) mask |= (7u << S_MU) | (1u << S_THU);
            block_reduce_loop(s, mask);
        }
        MD_TRACE(bid == 0 && tid == 0, 4);
        const bool last = publish_and_ticket<LOOP_BLOCK>(s, A.partials, sc);
        MD_TRACE(bid == 0 && tid == 0, 5);
        if (last) last_block_finalize<LOOP_BLOCK, true>(A.partials, sc, pr, FIN_STEP | (multi ? FIN_P2P : 0), A.peers);
        if (tid == 0) {
            const unsigned long long fin_seq0 = ctl.fin_seq;
            while (ld_acquire_gpu(&sc->fin_seq) <= fin_seq0) { }
            t_acc[3] += gtime() - tq;
            t_acc[4] += 1ull;
        }
        __syncthreads();
        MD_TRACE(bid == 0 && tid == 0, 6);
    }
    if (USMEM && loaded) {
        // the loop is over (rebuild, end of the batch, error or max_steps): the velocities go back to the planes — u, or v''
        // after the last step of a batch, exactly what the two-kernel step leaves there
        for (int p = 0; p < PS; ++p) {
            const int t = pair_of(p);
            if (t >= npairs) break;
            const int2 C = reinterpret_cast<const int2 *>(A.cntg)[t];
            if (((C.x | (2 * t + 1 < n ? C.y : 0)) & LOOP_GHOST_FLAG) != 0) continue;  // lives in the planes
            const double2 ux = su[(0 * PS + p) * LOOP_BLOCK + tid], uy = su[(1 * PS + p) * LOOP_BLOCK + tid],
                          uz = su[(2 * PS + p) * LOOP_BLOCK + tid];
            if (2 * t + 1 < n) {
                reinterpret_cast<double2 *>(a.vx)[t] = ux; reinterpret_cast<double2 *>(a.vy)[t] = uy;
                reinterpret_cast<double2 *>(a.vz)[t] = uz;
            } else {
                a.vx[2 * t] = ux.x; a.vy[2 * t] = uy.x; a.vz[2 * t] = uz.x;
            }
        }
    }
    if (bid == 0 && tid == 0 && t_acc[4]) {
        // (the last finalize of this launch is complete: nobody else writes these words)
#pragma unroll
        for (int k = 0; k < 4; ++k) sc->loop_ns[k] += t_acc[k];
        sc->loop_steps += t_acc[4];
    }
}

// pairs with a ghost partner (second pass of the loop's force phase)
__global__ void k_flag_bnd_pairs(int npairs, int n, const int *__restrict__ cntg, int *__restrict__ flag)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= npairs) return;
    const int c = cntg[2 * t] | (2 * t + 1 < n ? cntg[2 * t + 1] : 0);
    flag[t] = (c & LOOP_GHOST_FLAG) ? 1 : 0;
}

// the loop's copy of the list counts on one GPU is nbr_cnt itself; on the multi-GPU path the builder ORs in LOOP_GHOST_FLAG
__global__ void k_counts_with_ghost_flag(int n, const int *__restrict__ nbr_cnt, const int *__restrict__ has_ghost,
                                         int *__restrict__ cntg)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cntg[i] = nbr_cnt[i] | ((has_ghost[i] && nbr_cnt[i] > 0) ? LOOP_GHOST_FLAG : 0);
}

}  // namespace md

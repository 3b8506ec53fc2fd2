// moldyn_b200.cu — C ABI (include/moldyn_b200.h) over the sm_100a kernels in md_kernels.cuh.
//
// Host side of the step loop: owns the device-resident State, orchestrates list rebuilds, and runs the steady-state
// steps without host involvement — dilute systems inside ONE persistent cooperative kernel (md_loop.cuh), dense systems as
// pre-enqueued CUDA graphs of guarded {k_kick_drift; k_force} steps.  The host looks at the device only when the neighbour
// list has to be rebuilt or the batch ends.
#include "md_kernels.cuh"

#include <dlfcn.h>
#include <nccl.h>  // types only: the library is dlopen()ed on first multi-GPU use (see nccl_api below)

#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/moldyn_b200.h"

using namespace md;

// NCCL is bound at run time, not link time: single-GPU users never load it, and inside a process that has already
// imported PyTorch the dlopen() resolves to the libnccl.so.2 that is already mapped (two different NCCL builds in one
// process do not mix).
namespace nccl_api {
#define MD_NCCL_SYMBOLS(X)                                                                                   \
    X(ncclGetUniqueId) X(ncclCommInitRank) X(ncclCommDestroy) X(ncclSend) X(ncclRecv) X(ncclGroupStart)       \
    X(ncclGroupEnd) X(ncclAllGather) X(ncclGetErrorString)
#define X(name) decltype(&::name) name = nullptr;
MD_NCCL_SYMBOLS(X)
#undef X
const char *load()
{
    static const char *err = nullptr;
    static bool done = false;
    if (done) return err;
    done = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return err = "libnccl.so.2 not found (needed only for multi-GPU runs)";
#define X(name)                                         \
    name = (decltype(name))dlsym(h, #name);             \
    if (!name) return err = "libnccl lacks " #name;
    MD_NCCL_SYMBOLS(X)
#undef X
    return nullptr;
}
}  // namespace nccl_api
#define ncclGetUniqueId nccl_api::ncclGetUniqueId
#define ncclCommInitRank nccl_api::ncclCommInitRank
#define ncclCommDestroy nccl_api::ncclCommDestroy
#define ncclSend nccl_api::ncclSend
#define ncclRecv nccl_api::ncclRecv
#define ncclGroupStart nccl_api::ncclGroupStart
#define ncclGroupEnd nccl_api::ncclGroupEnd
#define ncclAllGather nccl_api::ncclAllGather
#define ncclGetErrorString nccl_api::ncclGetErrorString

namespace {

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct md_ctx {
    md_config cfg{};
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    // potential (PotentialsDatabase default: potential.rs:95-101)
    double sigma = 0.3418, eps = 1.712, r_cut = 0.0, u_cut = 0.0;
    double skin = 0.0;

    // state (n = global atom count; n_own/n_ghost = this rank's owned and halo atoms, n_own == n on one GPU)
    int64_t n = 0, n_own = 0, n_ghost = 0;
    int npad = 0;
    double mass = 0.0;
    bool has_state = false;
    // several particle types (md_multi.cuh / md_multi.inc): State.particles flattened type by type
    struct Multi {
        bool on = false;
        int T = 1;
        int mode = MD_CROSS_REFERENCE;
        int64_t start[MULTI_MAX_TYPES + 1] = {0};
        double mass[MULTI_MAX_TYPES] = {0};
        std::map<std::pair<int, int>, std::array<double, 4>> pairs;  // PotentialsDatabase entries other than (0, 0)
        double rc_max = 0.0;  // largest r_cut of the T x T table: the list radius
        double disp = 0.0;    // bound of any atom's displacement since the list build
        MultiTable *d_tab = nullptr, *h_tab = nullptr;  // h_*: pinned
        MultiWork *d_work = nullptr, *h_work = nullptr;
        double *d_partials = nullptr;
        int nblocks = 0;
    } multi;
    bool list_valid = false;
    bool force_valid = false;
    double sums_c = -1.0;  // half_dt_m the stored reduction was computed with
    bool sums_nh = false;  // the stored reduction includes the sums of u = v + F c (Nose-Hoover's second psi update)

    Arrays cur{}, alt{};
    std::vector<DevBuf> owned;
    double *stage = nullptr;  // 3*npad doubles, transfer staging
    int *stage_i = nullptr;   // npad ints
    Scalars *d_sc = nullptr;
    Params *d_pr = nullptr;
    Scalars *h_sc = nullptr;  // pinned
    Params *h_pr = nullptr;   // pinned
    Params prm{};
    double *d_partials = nullptr;
    int partial_blocks = 0;
    int force_grid[3] = {1, 1, 1}, reduce_grid = 1;  // exact, fast dense, fast dilute
    // persistent step loop (md_loop.cuh)
    int *nbr_ghost = nullptr;                         // multi-GPU: per-atom "the list holds a ghost" flags (list build)
    int *cntg = nullptr;                              // multi-GPU: list counts | LOOP_GHOST_FLAG (the loop's copy)
    int *bnd_pairs = nullptr, *n_bnd = nullptr;       // multi-GPU: pairs with a ghost partner, compacted; their number (device)
    int loop_blocks_max = 0;                          // co-resident blocks of k_md_loop on this device
    bool loop_attr_set = false;
    double rebuild_host_ms = 0.0;                     // multi-GPU: wall time spent in list rebuilds (host clock, synchronised)
    long long epoch_start_step = 0, last_epoch_len = 128;  // list epochs in steps: sizes the chunk look-ahead
    // dense + FAST on one GPU: brick tiles (md_tile.cuh) — 16-bit brick-local lists, atom-major
    unsigned short *nbrT = nullptr;
    size_t nbrT_alloc = 0;
    int cap16 = 0;                                    // list slots per atom (multiple of 32)
    int *brick_order = nullptr;                       // block → brick: full bricks first, the partial ones fill the tail
    int brick_alloc = 0, nbricks = 0;
    int sh_cap = 0, own_cap = 0;                      // shell slots / brick atoms the tile kernels' shared memory is sized for
    bool tile_valid = false;                          // the last rebuild produced tile lists
    bool tile_disabled = false;                       // this State cannot use them (a coordinate outside the box, shell too large)
    bool dense = false;                               // mean listed partners >= 8 at the last rebuild
    bool use_q4 = false;                              // packed gather copy maintained (dense systems)
    double graph_hc = -1.0;                           // dt/(2m) baked into the captured force kernel

    // cells / lists
    Grid grid{};
    int *cell_of = nullptr, *cell_sorted = nullptr, *order = nullptr;
    int *cell_cnt = nullptr, *cell_start = nullptr, *block_sums = nullptr;
    int cell_alloc = 0;
    int *nbr = nullptr, *nbr_cnt = nullptr;
    size_t nbr_alloc = 0;

    // multi-GPU (chunk path): a chunk of guarded steps incl. the NCCL calls
    cudaGraph_t dist_graph = nullptr;
    cudaGraphExec_t dist_graph_exec = nullptr;
    // single-GPU chunk graphs, kept by what is baked into them (see ChunkKey): a rebuild swaps the two plane sets, so the
    // graphs of two consecutive list epochs alternate — two entries make re-capture + re-instantiation a once-only cost
    struct ChunkGraph {
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        std::vector<unsigned char> key;
        unsigned long long stamp = 0;
    } chunk_cache[2];
    unsigned long long chunk_stamp = 0;
    long long chunk_hits = 0, chunk_builds = 0;

    md_stats stats{};

    // multi-GPU slab decomposition (md_dist.inc)
    struct Dist {
        bool on = false;
        int rank = 0, nranks = 1, left = 0, right = 0;
        ncclComm_t comm = nullptr;
        double *sendbuf[2] = {nullptr, nullptr}, *recvbuf[2] = {nullptr, nullptr};  // [0] left, [1] right neighbour
        int *send_ids[2] = {nullptr, nullptr}, *recv_ids[2] = {nullptr, nullptr};
        size_t buf_cap = 0;
        int *flag = nullptr, *scan = nullptr, *scan_sums = nullptr, *idx_stay = nullptr;
        int *idx_send[2] = {nullptr, nullptr};
        int *ghost_idx[2] = {nullptr, nullptr};  // owned atoms (sorted indices) whose positions go to each neighbour
        int ghost_send[2] = {0, 0}, ghost_recv[2] = {0, 0};
        int *ghost_cell_of = nullptr, *ghost_order = nullptr, *ghost_start = nullptr;
        int *d_cnt = nullptr, *h_cnt = nullptr;
        double *all_sums = nullptr;
        int64_t migrated = 0;
        // peer-memory path (NVLink stores into the neighbours' HBM; buffers shared through CUDA IPC)
        bool p2p = false;
        double *slab = nullptr;       // x, y, z of both plane sets + both q4 copies in ONE allocation (one IPC handle)
        int64_t slab_npad = 0;
        Mail *mail = nullptr;         // this rank's mailbox
        Peers peers{};                // every rank's mailbox as mapped here
        Peers *peers_dev = nullptr;   // the same, in device memory (kernel argument)
        double *peer_slab[2] = {nullptr, nullptr};  // left / right neighbour's slab as mapped here
        int64_t peer_npad[2] = {0, 0};
        int peer_base[2] = {0, 0}, peer_half[2] = {0, 0};  // where our face atoms land in the neighbour's planes
        int peer_own[2] = {0, 0};                          // the neighbours' owned-atom counts (ghost pulls of the step loop)
        unsigned int *push_ticket = nullptr;
        char *gather_buf = nullptr;   // small persistent device buffer for host-level all-gathers
        char *up_buf = nullptr;       // upload staging arena (kept between uploads)
        size_t up_bytes = 0;
        void *ipc_opened[2 * MAX_PEERS] = {};
        int n_ipc_opened = 0;
    } dist;

    // per-kernel CUDA-event timing (md_time_kernels)
    bool timing = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    double t_ms[4] = {0.0, 0.0, 0.0, 0.0};   // kick_drift / phase A, force / phase B + tail, rebuild, loop synchronisation
    int64_t t_cnt[4] = {0, 0, 0, 0};

    int fail(int code, const char *fmt, ...)
    {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        err = buf;
        return code;
    }
};

#define CK(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess)                                                                        \
            return ctx->fail(MD_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, \
                             __LINE__);                                                               \
    } while (0)

#define TRY(expr)              \
    do {                       \
        int rc_ = (expr);      \
        if (rc_ != MD_OK) return rc_; \
    } while (0)

namespace {

inline int blocks_for(int64_t n, int block) { return (int)((n + block - 1) / block); }

// Host restatement of Potential::get_potential_and_force (potential.rs:57-70) for the scalar entry points.
void lj_host(double sigma, double eps, double r_cut, double u_cut, double r, double *u, double *f)
{
    if (r > r_cut) {
        *u = 0.0;
        *f = 0.0;
        return;
    }
    volatile double sigma_r = sigma / r;
    volatile double x2 = sigma_r * sigma_r;
    volatile double x4 = x2 * x2;
    volatile double s6 = x2 * x4;
    volatile double s12 = s6 * s6;
    volatile double d = s12 - s6;
    volatile double a = 4.0 * eps;
    volatile double b = a * d;
    *u = b - u_cut;
    volatile double e = 24.0 * eps;
    volatile double g = e / r;
    volatile double h = 2.0 * s12;
    volatile double k = s6 - h;
    *f = g * k;
}

template <typename T>
int dev_alloc(md_ctx *ctx, T **out, size_t count)
{
    void *p = nullptr;
    size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    CK(cudaMalloc(&p, bytes));
    CK(cudaMemsetAsync(p, 0, bytes, ctx->stream));
    ctx->owned.push_back({p, bytes});
    *out = (T *)p;
    return MD_OK;
}

void dev_free(md_ctx *ctx, void *p)
{
    if (!p) return;
    for (auto &b : ctx->owned)
        if (b.p == p) {
            cudaFree(p);
            b.p = nullptr;
        }
}

int alloc_arrays(md_ctx *ctx, Arrays *a, int npad)
{
    TRY(dev_alloc(ctx, &a->x, npad)); TRY(dev_alloc(ctx, &a->y, npad)); TRY(dev_alloc(ctx, &a->z, npad));
    TRY(dev_alloc(ctx, &a->vx, npad)); TRY(dev_alloc(ctx, &a->vy, npad)); TRY(dev_alloc(ctx, &a->vz, npad));
    TRY(dev_alloc(ctx, &a->fx, npad)); TRY(dev_alloc(ctx, &a->fy, npad)); TRY(dev_alloc(ctx, &a->fz, npad));
    TRY(dev_alloc(ctx, &a->u, npad)); TRY(dev_alloc(ctx, &a->w, npad));
    TRY(dev_alloc(ctx, &a->id, npad));
    TRY(dev_alloc(ctx, &a->q4, npad));
    return MD_OK;
}

void free_arrays(md_ctx *ctx, Arrays *a)
{
    dev_free(ctx, a->x); dev_free(ctx, a->y); dev_free(ctx, a->z);
    dev_free(ctx, a->vx); dev_free(ctx, a->vy); dev_free(ctx, a->vz);
    dev_free(ctx, a->fx); dev_free(ctx, a->fy); dev_free(ctx, a->fz);
    dev_free(ctx, a->u); dev_free(ctx, a->w); dev_free(ctx, a->id); dev_free(ctx, a->q4);
    *a = Arrays{};
}

void drop_graph(md_ctx *ctx)
{
    if (ctx->dist_graph_exec) cudaGraphExecDestroy(ctx->dist_graph_exec);
    if (ctx->dist_graph) cudaGraphDestroy(ctx->dist_graph);
    ctx->dist_graph_exec = nullptr;
    ctx->dist_graph = nullptr;
}

int push_params(md_ctx *ctx)
{
    *ctx->h_pr = ctx->prm;
    CK(cudaMemcpyAsync(ctx->d_pr, ctx->h_pr, sizeof(Params), cudaMemcpyHostToDevice, ctx->stream));
    return MD_OK;
}

int pull_scalars(md_ctx *ctx)
{
    CK(cudaMemcpyAsync(ctx->h_sc, ctx->d_sc, sizeof(Scalars), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return MD_OK;
}

// radius the Verlet lists and the cell grid are built for (+ skin): the potential's r_cut, or the largest one of the table
double list_cut(const md_ctx *ctx) { return ctx->multi.on ? ctx->multi.rc_max : ctx->r_cut; }

void fill_potential_params(md_ctx *ctx)
{
    ctx->prm.sigma = ctx->sigma;
    ctx->prm.eps = ctx->eps;
    ctx->prm.r_cut = ctx->r_cut;
    ctx->prm.u_cut = ctx->u_cut;
    ctx->prm.r_list = list_cut(ctx) + ctx->skin;
    ctx->prm.mass = ctx->mass;
    ctx->prm.n = ctx->n;
}

// Skin: user value, else 0.075·r_cut for dense systems and r_cut for dilute ones where extra
// list entries are nearly free but every rebuild costs several steps of HBM traffic.
double choose_skin(const md_ctx *ctx, const double box[3])
{
    if (ctx->cfg.skin > 0.0) return ctx->cfg.skin;
    double volume = box[0] * box[1] * box[2];
    double rho = (double)ctx->n / volume;
    const double r_cut = list_cut(ctx);
    double in_cut = rho * 4.18879020478639 * r_cut * r_cut * r_cut;
    // dense: 0.075 r_cut — measured flat between 0.067 and 0.084 r_cut on C5 now that a rebuild costs 0.75 ms (was 0.1 r_cut)
    double skin = in_cut > 8.0 ? 0.075 * r_cut : r_cut;
    double min_box = std::min(box[0], std::min(box[1], box[2]));
    // keep r_list below half the smallest box edge when that is possible (single-image list semantics)
    if (r_cut + skin > 0.5 * min_box) skin = std::max(0.0, 0.5 * min_box - r_cut) * 0.5;
    return skin;
}

int choose_grid(md_ctx *ctx, const double box[3])
{
    Grid g{};
    double r_list = list_cut(ctx) + ctx->skin;
    double volume = box[0] * box[1] * box[2];
    // dense systems: half-size cells and a 5x5x5 stencil (fewer candidates per list build, tighter sorted order)
    const double in_list = (double)ctx->n / volume * 4.18879020478639 * r_list * r_list * r_list;
    int nsub = ctx->cfg.cell_subdiv >= 2 ? 2 : (ctx->cfg.cell_subdiv == 1 ? 1 : (in_list > 40.0 ? 2 : 1));
    double k = ctx->cfg.cell_atoms > 0.0 ? ctx->cfg.cell_atoms : 1.0;
    double dilute_edge = std::cbrt(k * volume / (double)ctx->n);
    double edge = std::max(r_list / nsub * 1.02, dilute_edge);
    int64_t ncell = 1;
    for (int d = 0; d < 3; ++d) {
        int c = (int)std::floor(box[d] / edge);
        if (c < 1) c = 1;
        g.nc[d] = c;
        ncell *= c;
    }
    // Dense systems, FAST arithmetic, one GPU: brick order for the tile kernels (md_tile.cuh).  MOLDYN_B200_TILE=0 keeps the
    // per-thread Verlet path (A/B measurements); MOLDYN_B200_TILE_BZ fixes the z cells per brick (default: chosen below).
    static const int tile_env = [] { const char *e = std::getenv("MOLDYN_B200_TILE"); return e ? atoi(e) : 1; }();
    static const int bz_env = [] { const char *e = std::getenv("MOLDYN_B200_TILE_BZ"); return e ? atoi(e) : 0; }();
    g.brick = 0;
    if (tile_env != 0 && !ctx->tile_disabled && !ctx->dist.on && ctx->cfg.force_mode == MD_FORCE_FAST && nsub == 2 &&
        g.nc[0] >= TILE_MIN_CELLS && g.nc[1] >= TILE_MIN_CELLS && g.nc[2] >= TILE_MIN_CELLS && ctx->n >= 128) {
        g.brick = 1;
        g.nbx = (g.nc[0] + 3) / 4;
        g.nby = (g.nc[1] + 3) / 4;
        // z cells per brick: the force kernel runs two blocks per SM and every brick costs about the same, so the kernel
        // takes ceil(bricks / (2 SMs)) rounds of one brick each.  A brick costs its own cells' pair work plus the staging of
        // its (bz + 4) x 8 x 8 shell plus a fixed part (weights from C5 on B200: 0.92 us per own cell of 16 columns, 0.01 us
        // per staged cell, 2 us fixed); shorter bricks stage more per atom but can fill the last round.
        int bz = bz_env;
        if (bz <= 0) {
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
            double best = 0.0;
            for (int cand = 4; cand >= 2; --cand) {
                const long long bricks = (long long)g.nbx * g.nby * ((g.nc[2] + cand - 1) / cand);
                const double rounds = (double)((bricks + 2 * sms - 1) / (2 * sms));
                const double cost = rounds * (0.92 * 16.0 * cand + 0.01 * 64.0 * (cand + 4) + 2.0);
                if (bz <= 0 || cost < best * 0.97) { bz = cand; best = cost; }
            }
        }
        g.bz = std::max(1, std::min(bz, g.nc[2] - 4));
        g.nbz = (g.nc[2] + g.bz - 1) / g.bz;
        ncell = (int64_t)g.nbx * g.nby * 16 * g.nc[2];
    }
    if (ncell > (int64_t)1 << 30) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "cell grid too large");
    g.nsub = nsub;
    g.ncell = (int)ncell;
    g.cap = ctx->grid.cap;
    g.npad = ctx->npad;
    ctx->grid = g;
    if (g.ncell + 1 > ctx->cell_alloc) {
        dev_free(ctx, ctx->cell_cnt); dev_free(ctx, ctx->cell_start); dev_free(ctx, ctx->block_sums);
        ctx->cell_alloc = g.ncell + 1 + g.ncell / 8;
        TRY(dev_alloc(ctx, &ctx->cell_cnt, ctx->cell_alloc));
        TRY(dev_alloc(ctx, &ctx->cell_start, ctx->cell_alloc));
        TRY(dev_alloc(ctx, &ctx->block_sums, blocks_for(ctx->cell_alloc, SCAN_BLOCK) + 1));
    }
    if (g.brick) {
        // block → brick: by the number of real cells, descending (stable): the partial bricks at the upper faces are the
        // cheap ones and fill the tail of the grid
        ctx->nbricks = g.nbx * g.nby * g.nbz;
        std::vector<int> order((size_t)ctx->nbricks);
        std::vector<long long> weight((size_t)ctx->nbricks);
        for (int id = 0; id < ctx->nbricks; ++id) {
            const int bzc = id % g.nbz, bxy = id / g.nbz, by = bxy % g.nby, bx = bxy / g.nby;
            const long long wx = std::min(4, g.nc[0] - 4 * bx), wy = std::min(4, g.nc[1] - 4 * by),
                            wz = std::min(g.bz, g.nc[2] - g.bz * bzc);
            order[(size_t)id] = id;
            weight[(size_t)id] = wx * wy * wz;
        }
        std::stable_sort(order.begin(), order.end(), [&](int p, int q) { return weight[(size_t)p] > weight[(size_t)q]; });
        if (ctx->nbricks > ctx->brick_alloc) {
            dev_free(ctx, ctx->brick_order);
            ctx->brick_alloc = ctx->nbricks + ctx->nbricks / 8;
            TRY(dev_alloc(ctx, &ctx->brick_order, ctx->brick_alloc));
        }
        CK(cudaMemcpyAsync(ctx->brick_order, order.data(), sizeof(int) * (size_t)ctx->nbricks, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));  // (order is a local)
    }
    return MD_OK;
}

bool grid_still_valid(const md_ctx *ctx, const double box[3])
{
    double r_list = list_cut(ctx) + ctx->skin;
    for (int d = 0; d < 3; ++d) {
        int w = 2 * ctx->grid.nsub + 1;
        if (ctx->grid.nc[d] >= w && box[d] / ctx->grid.nc[d] * ctx->grid.nsub < r_list) return false;
    }
    return true;
}

int ensure_nbr_capacity(md_ctx *ctx, int cap)
{
    size_t need = (size_t)cap * (size_t)ctx->npad;
    if (need > ctx->nbr_alloc) {
        dev_free(ctx, ctx->nbr);
        ctx->nbr = nullptr;
        ctx->nbr_alloc = 0;
        TRY(dev_alloc(ctx, &ctx->nbr, need));
        ctx->nbr_alloc = need;
    }
    ctx->grid.cap = cap;
    drop_graph(ctx);
    return MD_OK;
}

// Largest double t with sqrt(t) <= r (IEEE sqrt is correctly rounded and monotonic): turns the reference's
// `norm(r) > r_cut` test into a comparison of squares with the identical outcome for every input.
double sqrt_threshold(double r)
{
    double t = r * r;
    while (std::sqrt(t) > r) t = std::nextafter(t, 0.0);
    while (std::sqrt(std::nextafter(t, INFINITY)) <= r) t = std::nextafter(t, INFINITY);
    return t;
}

ForceConsts force_consts(const md_ctx *ctx)
{
    ForceConsts fc;
    fc.sigma = ctx->sigma;
    fc.sigma2 = ctx->sigma * ctx->sigma;
    fc.eps4 = 4.0 * ctx->eps;    // == __dmul_rn(4.0, eps): the product the reference forms (potential.rs:67)
    fc.eps24 = 24.0 * ctx->eps;  // potential.rs:68
    fc.r_cut = ctx->r_cut;
    fc.rc2 = ctx->r_cut * ctx->r_cut;
    fc.u_cut = ctx->u_cut;
    {
        const double s2 = ctx->sigma * ctx->sigma, s6 = s2 * s2 * s2;
        fc.c6 = 24.0 * ctx->eps * s6;
        fc.c12 = 48.0 * ctx->eps * s6 * s6;
        fc.d6 = 4.0 * ctx->eps * s6;
        fc.d12 = 4.0 * ctx->eps * s6 * s6;
    }
    fc.hc = ctx->prm.half_dt_m;
    fc.mass = ctx->mass;
    return fc;
}

int refresh_q4(md_ctx *ctx)
{
    if (!ctx->use_q4) return MD_OK;
    const int n = (int)(ctx->n_own + ctx->n_ghost);
    k_pack_q4<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(n, ctx->cur);
    ctx->stats.kernel_launches += 1;
    return MD_OK;
}

HaloPush dist_halo_push_args(md_ctx *ctx);
void dist_ghost_pull_args(md_ctx *ctx, LoopArgs *A);
int build_tile_lists(md_ctx *ctx, bool *fallback);

// K1 + K2, host-orchestrated (rare: every O(10-100) steps).  Positions must be consistent with the box
// (no pending barostat scaling).
int rebuild_lists(md_ctx *ctx)
{
    cudaStream_t st = ctx->stream;
    const int n = (int)ctx->n;
    TRY(pull_scalars(ctx));
    double box[3] = {ctx->h_sc->box[0], ctx->h_sc->box[1], ctx->h_sc->box[2]};
    for (int d = 0; d < 3; ++d)
        if (!(box[d] > 0.0) || !std::isfinite(box[d])) return ctx->fail(MD_ERR_NONFINITE, "boundary box is not finite/positive");
    if (ctx->grid.ncell == 0 || !grid_still_valid(ctx, box) || !ctx->list_valid) TRY(choose_grid(ctx, box));
    Grid &g = ctx->grid;

    CK(cudaMemsetAsync(ctx->cell_cnt, 0, sizeof(int) * (g.ncell + 1), st));
    k_cell_count<<<blocks_for(n, 256), 256, 0, st>>>(n, ctx->cur.x, ctx->cur.y, ctx->cur.z, ctx->d_sc, g,
                                                      ctx->cell_of, ctx->cell_cnt);
    int sb = blocks_for(g.ncell, SCAN_BLOCK);
    k_scan_block<<<sb, SCAN_BLOCK, 0, st>>>(g.ncell, ctx->cell_cnt, ctx->cell_start, ctx->block_sums);
    k_scan_sums<<<1, SCAN_BLOCK, 0, st>>>(sb, ctx->block_sums);
    k_scan_add<<<sb, SCAN_BLOCK, 0, st>>>(g.ncell, ctx->cell_start, ctx->block_sums, n);
    CK(cudaMemsetAsync(ctx->cell_cnt, 0, sizeof(int) * (g.ncell + 1), st));
    k_scatter<<<blocks_for(n, 256), 256, 0, st>>>(n, ctx->cell_of, ctx->cell_start, ctx->cell_cnt, ctx->order);
    k_sort_cells<<<blocks_for(g.ncell, 128), 128, 0, st>>>(g.ncell, ctx->cell_start, ctx->cur.id, ctx->order);
    k_reorder<<<blocks_for(n, 256), 256, 0, st>>>(n, ctx->order, ctx->cell_of, ctx->cur, ctx->alt,
                                                   ctx->cell_sorted);
    std::swap(ctx->cur, ctx->alt);
    ctx->stats.kernel_launches += 7;

    const double r2_list = sqrt_threshold(ctx->prm.r_list);
    ctx->tile_valid = false;
    if (g.brick) {
        // dense systems: brick-local 16-bit lists for the tile force kernel (md_tile.cuh)
        bool fallback = false;
        TRY(build_tile_lists(ctx, &fallback));
        if (fallback) {
            // a coordinate outside the box, or a shell that does not fit in shared memory: canonical order, per-thread lists
            ctx->tile_disabled = true;
            ctx->list_valid = false;
            return rebuild_lists(ctx);
        }
        ctx->tile_valid = true;
    }
    for (int attempt = 0; !g.brick && attempt < 4; ++attempt) {
        k_reset_list_stats<<<1, 1, 0, st>>>(ctx->d_sc);
        // image shift per cell run instead of per candidate when the box is wide enough in cells (see k_build_list)
        const int need_cells = 2 * g.nsub + 3;
        const bool shift = g.nc[0] >= need_cells && g.nc[1] >= need_cells && g.nc[2] >= need_cells;
#define LAUNCH_BUILD(SORT, SHIFT)                                                                                     \
    k_build_list<SORT, SHIFT><<<blocks_for(n, 128), 128, 0, st>>>(n, ctx->cur, ctx->cell_sorted, ctx->cell_start,     \
                                                                 ctx->d_sc, g, ctx->prm.r_list, r2_list, ctx->nbr,   \
                                                                 ctx->nbr_cnt)
        if (ctx->cfg.force_mode == MD_FORCE_EXACT) {
            if (shift) LAUNCH_BUILD(true, true); else LAUNCH_BUILD(true, false);
        } else {
            if (shift) LAUNCH_BUILD(false, true); else LAUNCH_BUILD(false, false);
        }
#undef LAUNCH_BUILD
        ctx->stats.kernel_launches += 2;
        CK(cudaGetLastError());
        TRY(pull_scalars(ctx));
        if (!ctx->h_sc->nbr_overflow) break;
        if (attempt == 3) return ctx->fail(MD_ERR_NEIGHBOUR_OVERFLOW, "neighbour list overflow (max %d)", ctx->h_sc->nbr_max);
        int cap = ((int)(ctx->h_sc->nbr_max * 1.25) + 8 + 7) / 8 * 8;
        TRY(ensure_nbr_capacity(ctx, cap));
    }
    k_after_rebuild<<<1, 1, 0, st>>>(ctx->d_sc);
    ctx->stats.kernel_launches += 1;
    CK(cudaGetLastError());
    ctx->stats.rebuilds += 1;
    if (ctx->stats.steps > ctx->epoch_start_step) ctx->last_epoch_len = ctx->stats.steps - ctx->epoch_start_step;
    ctx->epoch_start_step = ctx->stats.steps;
    ctx->stats.nbr_max = ctx->h_sc->nbr_max;
    ctx->stats.nbr_mean = (double)ctx->h_sc->nbr_total / (double)ctx->n;
    ctx->dense = ctx->stats.nbr_mean >= 8.0 && n >= 128;  // (the dense kernel's masked lanes need a foreign warp's atom)
    ctx->use_q4 = ctx->dense && ctx->cfg.force_mode != MD_FORCE_EXACT && !ctx->tile_valid;
    TRY(refresh_q4(ctx));
    ctx->list_valid = true;
    return MD_OK;
}

// K2 in tile form: shell sizes first (they size the kernels' shared memory), then the brick-local lists.
int build_tile_lists(md_ctx *ctx, bool *fallback)
{
    cudaStream_t st = ctx->stream;
    const Grid &g = ctx->grid;
    *fallback = false;
    k_tile_reset<<<1, 1, 0, st>>>(ctx->d_sc);
    k_tile_measure<<<ctx->nbricks, 64, 0, st>>>(g, ctx->cell_start, ctx->d_sc);
    ctx->stats.kernel_launches += 2;
    CK(cudaGetLastError());
    TRY(pull_scalars(ctx));
    if (ctx->h_sc->out_of_box) { *fallback = true; return MD_OK; }
    ctx->sh_cap = (ctx->h_sc->tile_shell_max + 1 + 63) / 64 * 64;  // (+1: the padding slot, md_tile.cuh)
    ctx->own_cap = (ctx->h_sc->tile_own_max + 31) / 32 * 32;
    int optin = 0;
    CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ctx->device));
    const size_t smem_force = ((size_t)3 * ctx->sh_cap + (size_t)5 * ctx->own_cap) * sizeof(double);
    const size_t smem_build = (size_t)ctx->sh_cap * sizeof(float4) + (size_t)TILE_BLOCK * TILE_BUF * sizeof(unsigned short);
    if (ctx->sh_cap > 65535 || smem_force + 16 * 1024 > (size_t)optin) { *fallback = true; return MD_OK; }
    if (ctx->nbricks > ctx->partial_blocks) {  // one slot of partial sums per brick
        dev_free(ctx, ctx->d_partials);
        ctx->partial_blocks = ctx->nbricks + ctx->nbricks / 8;
        TRY(dev_alloc(ctx, &ctx->d_partials, (size_t)ctx->partial_blocks * NSUM));
    }
    CK(cudaFuncSetAttribute(k_force_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_force));
    CK(cudaFuncSetAttribute(k_build_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_build));
    CK(cudaFuncSetAttribute(k_tile_expand, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)ctx->sh_cap * sizeof(int))));
    const double r2_list = sqrt_threshold(ctx->prm.r_list);
    if (ctx->cap16 == 0) {
        const double volume = ctx->h_sc->box[0] * ctx->h_sc->box[1] * ctx->h_sc->box[2];
        const double expect = (double)ctx->n / volume * 4.18879020478639 * ctx->prm.r_list * ctx->prm.r_list * ctx->prm.r_list;
        ctx->cap16 = std::max(64, ((int)(expect * 1.3) + 16 + 31) / 32 * 32);  // (>= 64: the pair loop's first two trips)
    }
    for (int attempt = 0; attempt < 4; ++attempt) {
        const size_t need = (size_t)ctx->cap16 * (size_t)ctx->npad;
        if (need > ctx->nbrT_alloc) {
            dev_free(ctx, ctx->nbrT);
            ctx->nbrT = nullptr;
            ctx->nbrT_alloc = 0;
            TRY(dev_alloc(ctx, &ctx->nbrT, need));
            ctx->nbrT_alloc = need;
        }
        k_reset_list_stats<<<1, 1, 0, st>>>(ctx->d_sc);
        k_build_tile<<<ctx->nbricks, TILE_BLOCK, smem_build, st>>>(g, ctx->cur, ctx->cell_start, ctx->cell_sorted, ctx->d_sc,
                                                                   r2_list, ctx->nbrT, ctx->cap16, ctx->nbr_cnt,
                                                                   ctx->brick_order, ctx->sh_cap);
        ctx->stats.kernel_launches += 2;
        CK(cudaGetLastError());
        TRY(pull_scalars(ctx));
        if (!ctx->h_sc->nbr_overflow) break;
        if (attempt == 3) return ctx->fail(MD_ERR_NEIGHBOUR_OVERFLOW, "neighbour list overflow (max %d)", ctx->h_sc->nbr_max);
        ctx->cap16 = std::max(64, ((int)(ctx->h_sc->nbr_max * 1.15) + 8 + 31) / 32 * 32);
    }
    return MD_OK;
}

int launch_kick_drift(md_ctx *ctx, int guarded = 0, const HaloPush *push = nullptr)
{
    const int n = (int)ctx->n_own;
    const int blocks = std::max(1, blocks_for((n + 1) / 2, 256));
    k_kick_drift<<<blocks, 256, 0, ctx->stream>>>(n, ctx->cur, ctx->d_sc, ctx->d_pr, guarded, ctx->use_q4 ? 1 : 0,
                                                  push ? *push : HaloPush{});
    return MD_OK;
}

// Dilute systems run their steps inside the persistent loop (md_loop.cuh); dense ones as graph chunks of the two-kernel step.
bool loop_wanted(const md_ctx *ctx)
{
    if (ctx->dense || ctx->cfg.loop_mode == MD_LOOP_CHUNK) return false;
    if (ctx->dist.on) return ctx->dist.p2p;
    // one GPU: with a few partners per atom the two-kernel step's force kernel (two atoms per thread, operands of the next
    // pair prefetched) walks its lists faster than the loop's force phase — measured on C1's steady state (4.9 listed
    // partners: 14.5 vs 27 us/step).  The loop wins where fixed latencies dominate: small systems (C2: a tie at 0.55
    // partners), nearly empty lists (C3 from the lattice: 22.2 vs 28-30 us/step) and states beyond L2 (8*10^6 atoms: 174 vs
    // 197).  For 10^5..2.6*10^6 atoms with populated lists (C3's collisional steady state, 0.55 partners) the two-kernel
    // step is 5-10 % ahead on three different boxes (35 vs 37-39 us/step, profiles/r02_call_{i,o,r}_*.txt).
    const bool mid_size = ctx->n >= 131072 && ctx->n <= 2600000;
    return ctx->stats.nbr_mean < (mid_size ? 0.25 : 2.0);
}

int launch_force(md_ctx *ctx, bool kick, int guarded = 0)
{
    const int n = (int)ctx->n;
    const ForceConsts fc = force_consts(ctx);
    const int flags = (kick ? 1 : 0) | (guarded ? 4 : 0);
#define LAUNCH_FORCE(E, M, GRID)                                                                                      \
    k_force<E, M><<<GRID, FORCE_BLOCK, 0, ctx->stream>>>(n, ctx->cur, ctx->nbr, ctx->nbr_cnt, ctx->npad, ctx->grid.cap, \
                                                         ctx->d_partials, ctx->d_sc, ctx->d_pr, flags, fc, nullptr)
    if (ctx->cfg.force_mode == MD_FORCE_EXACT) LAUNCH_FORCE(true, false, ctx->force_grid[0]);
    else if (ctx->tile_valid) {
        const size_t smem = ((size_t)3 * ctx->sh_cap + (size_t)5 * ctx->own_cap) * sizeof(double);
        k_force_tile<<<ctx->nbricks, TILE_BLOCK, smem, ctx->stream>>>(ctx->grid, ctx->cur, ctx->cell_start, ctx->nbrT, ctx->cap16,
                                                                      ctx->nbr_cnt, ctx->d_partials, ctx->d_sc, ctx->d_pr, flags,
                                                                      fc, ctx->brick_order, ctx->sh_cap);
    } else if (ctx->dense) LAUNCH_FORCE(false, true, ctx->force_grid[1]);
    else LAUNCH_FORCE(false, false, ctx->force_grid[2]);
#undef LAUNCH_FORCE
    return MD_OK;
}

int launch_reduce(md_ctx *ctx)
{
    const int n = (int)ctx->n;
    k_reduce_state<<<ctx->reduce_grid, RED_BLOCK, 0, ctx->stream>>>(n, ctx->cur, ctx->d_partials, ctx->d_sc, ctx->d_pr, 0);
    ctx->stats.kernel_launches += 1;
    return MD_OK;
}

constexpr size_t LOOP_SMEM_PER_PAIR = 3 * LOOP_BLOCK * sizeof(double2);  // ux, uy, uz of one pair per thread
constexpr size_t LOOP_SMEM_MAX = LOOP_MAX_PAIRS * LOOP_SMEM_PER_PAIR;

// Persistent grids: resident blocks per SM (occupancy API) × SM count, capped by the work available.
int choose_grids(md_ctx *ctx)
{
    int sms = 0, occ[3] = {0, 0, 0}, occ_r = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], k_force<true, false>, FORCE_BLOCK, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], k_force<false, true>, FORCE_BLOCK, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], k_force<false, false>, FORCE_BLOCK, 0));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_r, k_reduce_state, RED_BLOCK, 0));
    const int pair_blocks = blocks_for((ctx->n + 1) / 2, FORCE_BLOCK);
    for (int k = 0; k < 3; ++k) ctx->force_grid[k] = std::max(1, std::min(pair_blocks, sms * std::max(occ[k], 1)));
    ctx->reduce_grid = std::max(1, std::min(blocks_for(ctx->n, RED_BLOCK), sms * std::max(occ_r, 1)));
    if (!ctx->loop_attr_set) {
        CK(cudaFuncSetAttribute(k_md_loop<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LOOP_SMEM_MAX));
        CK(cudaFuncSetAttribute(k_md_loop<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)LOOP_SMEM_MAX));
        int occ_l[4] = {0, 0, 0, 0};
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_l[0], k_md_loop<false, true>, LOOP_BLOCK, LOOP_SMEM_MAX));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_l[1], k_md_loop<true, true>, LOOP_BLOCK, LOOP_SMEM_MAX));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_l[2], k_md_loop<false, false>, LOOP_BLOCK, 0));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_l[3], k_md_loop<true, false>, LOOP_BLOCK, 0));
        int coop = 0;
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, ctx->device));
        const int occ_min = std::min(std::min(occ_l[0], occ_l[1]), std::min(occ_l[2], occ_l[3]));
        ctx->loop_blocks_max = coop ? sms * std::min(1, occ_min) : 0;  // one block per SM
        ctx->loop_attr_set = true;
    }
    return MD_OK;
}

// One launch of the persistent step loop: runs until the list has to be rebuilt, the batch ends or max_steps are done.
int launch_loop(md_ctx *ctx, long long max_steps)
{
    const int n = (int)ctx->n_own;
    const int npairs = (n + 1) / 2;
    // as many blocks as there are pairs of atoms to drift, at most one per SM: small systems synchronise a small grid
    const int grid = std::max(1, std::min(ctx->loop_blocks_max, blocks_for(npairs, LOOP_BLOCK)));
    const int P = std::max(1, blocks_for(npairs, (int64_t)grid * LOOP_BLOCK));
    // how many of a thread's pairs keep their velocities in shared memory: all of them when they fit (P <= LOOP_MAX_PAIRS),
    // else none.  MOLDYN_B200_LOOP_PSMEM=k keeps the first min(P, k) instead (A/B: shared memory against L1 capacity for
    // the gathers; larger systems partly resident).
    static const int ps_env = [] { const char *e = std::getenv("MOLDYN_B200_LOOP_PSMEM"); return e ? atoi(e) : -1; }();
    const int PS = ps_env >= 0 ? std::min(P, std::min(ps_env, LOOP_MAX_PAIRS)) : (P <= LOOP_MAX_PAIRS ? P : 0);
    const bool usmem = PS > 0;
    LoopArgs A{};
    A.n = n;
    A.npad = ctx->npad;
    A.cap = ctx->grid.cap;
    A.pairs_per_thread = P;
    A.pairs_in_smem = PS;
    A.a = ctx->cur;
    A.nbr = ctx->nbr;
    A.cntg = ctx->dist.on ? ctx->cntg : ctx->nbr_cnt;
    A.bnd_pairs = ctx->bnd_pairs;
    A.n_bnd = ctx->n_bnd;
    A.partials = ctx->d_partials;
    A.sc = ctx->d_sc;
    A.pr = ctx->d_pr;
    A.peers = ctx->dist.on ? ctx->dist.peers_dev : nullptr;
    if (ctx->dist.on) dist_ghost_pull_args(ctx, &A);
    A.max_steps = max_steps;
    A.fc = force_consts(ctx);
    void *args[] = {&A};
    const size_t smem = (size_t)PS * LOOP_SMEM_PER_PAIR;
    const bool exact = ctx->cfg.force_mode == MD_FORCE_EXACT;
    const void *fn = exact ? (usmem ? (const void *)k_md_loop<true, true> : (const void *)k_md_loop<true, false>)
                           : (usmem ? (const void *)k_md_loop<false, true> : (const void *)k_md_loop<false, false>);
    CK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(LOOP_BLOCK), args, smem, ctx->stream));
    ctx->stats.loop_launches += 1;
    return MD_OK;
}

// MD_LOOP_CHUNK: STEP_CHUNK guarded steps captured per list epoch; the host enqueues a few of them ahead.
constexpr int STEP_CHUNK = 16;

void drop_chunk_cache(md_ctx *ctx)
{
    for (auto &c : ctx->chunk_cache) {
        if (c.exec) cudaGraphExecDestroy(c.exec);
        if (c.graph) cudaGraphDestroy(c.graph);
        c.exec = nullptr;
        c.graph = nullptr;
        c.key.clear();
    }
}

// Everything launch_kick_drift / launch_force bake into a captured chunk: kernel variant selectors, grids and every
// argument.  Equal keys mean byte-identical launches, so a cached graph may be replayed whatever happened in between.
std::vector<unsigned char> chunk_key(const md_ctx *ctx)
{
    std::vector<unsigned char> k;
    auto put = [&k](const void *p, size_t n) { const unsigned char *b = (const unsigned char *)p; k.insert(k.end(), b, b + n); };
    auto put_ptr = [&put](const void *p) { put(&p, sizeof p); };
    auto put_i = [&put](long long v) { put(&v, sizeof v); };
    // Arrays hold only pointers; copy them field by field (no padding bytes in the key)
    for (const Arrays *a : {&ctx->cur, &ctx->alt})
        for (const void *q : {(const void *)a->x, (const void *)a->y, (const void *)a->z, (const void *)a->vx, (const void *)a->vy,
                              (const void *)a->vz, (const void *)a->fx, (const void *)a->fy, (const void *)a->fz,
                              (const void *)a->u, (const void *)a->w, (const void *)a->id, (const void *)a->q4})
            put_ptr(q);
    for (const void *q : {(const void *)ctx->nbr, (const void *)ctx->nbr_cnt, (const void *)ctx->d_partials,
                          (const void *)ctx->d_sc, (const void *)ctx->d_pr, (const void *)ctx->nbrT,
                          (const void *)ctx->cell_start, (const void *)ctx->brick_order})
        put_ptr(q);
    const ForceConsts fc = force_consts(ctx);
    for (double v : {fc.sigma, fc.sigma2, fc.eps4, fc.eps24, fc.r_cut, fc.rc2, fc.u_cut, fc.c6, fc.c12, fc.d6, fc.d12, fc.hc, fc.mass})
        put(&v, sizeof v);
    for (long long v : {(long long)ctx->n, (long long)ctx->n_own, (long long)ctx->npad, (long long)ctx->grid.cap,
                        (long long)ctx->cfg.force_mode, (long long)ctx->dense, (long long)ctx->use_q4,
                        (long long)ctx->force_grid[0], (long long)ctx->force_grid[1], (long long)ctx->force_grid[2],
                        (long long)ctx->tile_valid, (long long)ctx->cap16, (long long)ctx->sh_cap, (long long)ctx->own_cap,
                        (long long)ctx->nbricks, (long long)ctx->grid.nc[0], (long long)ctx->grid.nc[1],
                        (long long)ctx->grid.nc[2], (long long)ctx->grid.bz})
        put_i(v);
    return k;
}

int build_chunk_graph(md_ctx *ctx, md_ctx::ChunkGraph *slot)
{
    if (slot->exec) cudaGraphExecDestroy(slot->exec);
    if (slot->graph) cudaGraphDestroy(slot->graph);
    slot->exec = nullptr;
    slot->graph = nullptr;
    slot->key.clear();
    const int64_t launches = ctx->stats.kernel_launches;
    CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    int rc = MD_OK;
    for (int k = 0; k < STEP_CHUNK && rc == MD_OK; ++k) {
        rc = launch_kick_drift(ctx, 1);
        if (rc == MD_OK) rc = launch_force(ctx, true, 1);
    }
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &slot->graph);  // always leave capture mode
    ctx->stats.kernel_launches = launches;
    if (rc != MD_OK) return rc;
    if (e != cudaSuccess) return ctx->fail(MD_ERR_CUDA, "graph capture of the step chunk failed: %s", cudaGetErrorString(e));
    CK(cudaGraphInstantiate(&slot->exec, slot->graph, 0));
    slot->key = chunk_key(ctx);
    ctx->chunk_builds += 1;
    return MD_OK;
}

// the cached chunk graph for the current state of the context, built on a miss (evicting the older entry)
int get_chunk_graph(md_ctx *ctx, cudaGraphExec_t *exec)
{
    const std::vector<unsigned char> key = chunk_key(ctx);
    md_ctx::ChunkGraph *hit = nullptr, *victim = &ctx->chunk_cache[0];
    for (auto &c : ctx->chunk_cache) {
        if (c.exec && c.key == key) hit = &c;
        if (c.stamp < victim->stamp) victim = &c;
    }
    if (!hit) {
        TRY(build_chunk_graph(ctx, victim));
        hit = victim;
    } else {
        ctx->chunk_hits += 1;
    }
    hit->stamp = ++ctx->chunk_stamp;
    *exec = hit->exec;
    return MD_OK;
}

int flush_pending_scale(md_ctx *ctx)
{
    const int n = (int)ctx->n_own;
    k_scale_positions<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(n, ctx->cur, ctx->d_sc);
    k_clear_pending<<<1, 1, 0, ctx->stream>>>(ctx->d_sc);
    ctx->stats.kernel_launches += 2;
    TRY(refresh_q4(ctx));
    CK(cudaGetLastError());
    return MD_OK;
}

int check_ctx(md_ctx *ctx, bool need_state)
{
    if (!ctx) return MD_ERR_INVALID_ARGUMENT;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return ctx->fail(MD_ERR_CUDA, "cudaSetDevice(%d): %s", ctx->device, cudaGetErrorString(e));
    if (need_state && !ctx->has_state) return ctx->fail(MD_ERR_NO_STATE, "no state uploaded");
    return MD_OK;
}

int device_error(md_ctx *ctx)
{
    if (ctx->h_sc->error == MD_ERR_NONFINITE)
        return ctx->fail(MD_ERR_NONFINITE, "non-finite thermostat/barostat coefficient or displacement "
                                           "(temperature %.17g, pressure %.17g)", ctx->h_sc->temperature,
                         ctx->h_sc->pressure);
    if (ctx->h_sc->error) return ctx->fail(ctx->h_sc->error, "device-side error %d", ctx->h_sc->error);
    return MD_OK;
}

#include "md_dist.inc"
#include "md_multi.inc"

}  // namespace

// =====================================================================================================
extern "C" {

const char *md_version(void) { return "moldyn_b200 0.1 (sm_100a)"; }

const char *md_last_error(const md_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int md_lj_potential_and_force(double sigma, double eps, double r_cut, double u_cut, double r, double *potential,
                              double *force)
{
    if (!potential || !force) return MD_ERR_INVALID_ARGUMENT;
    lj_host(sigma, eps, r_cut, u_cut, r, potential, force);
    return MD_OK;
}

int md_lj_new(double sigma, double eps, double *r_cut, double *u_cut)
{
    if (!r_cut || !u_cut) return MD_ERR_INVALID_ARGUMENT;
    double rc = sigma * 2.5, u, f;  // potential.rs:28
    lj_host(sigma, eps, rc, 0.0, rc, &u, &f);
    *r_cut = rc;
    *u_cut = u;
    return MD_OK;
}

int md_create(const md_config *cfg, md_ctx **out)
{
    if (!out) return MD_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    md_ctx *ctx = new md_ctx();
    if (cfg) ctx->cfg = *cfg;
    if (const char *e = std::getenv("MOLDYN_B200_LOOP"))  // profiling aid: ncu cannot see kernels of conditional graphs
    {
        if (!strcmp(e, "host")) ctx->cfg.loop_mode = MD_LOOP_HOST;
        if (!strcmp(e, "chunk")) ctx->cfg.loop_mode = MD_LOOP_CHUNK;
    }
    ctx->device = ctx->cfg.device;
    auto bail = [&](cudaError_t e, const char *what) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(e) +
                         " (moldyn_b200 has no CPU fallback; a CUDA device is required)";
        delete ctx;
        return (int)MD_ERR_CUDA;
    };
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return bail(e, "cudaGetDeviceCount");
    if (count == 0) return bail(cudaErrorNoDevice, "cudaGetDeviceCount");
    if (ctx->device < 0 || ctx->device >= count) {
        g_create_error = "device ordinal out of range";
        delete ctx;
        return MD_ERR_INVALID_ARGUMENT;
    }
    if ((e = cudaSetDevice(ctx->device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess)
        return bail(e, "cudaStreamCreate");
    if ((e = cudaMalloc((void **)&ctx->d_sc, sizeof(Scalars))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMalloc((void **)&ctx->d_pr, sizeof(Params))) != cudaSuccess) return bail(e, "cudaMalloc");
    if ((e = cudaMallocHost((void **)&ctx->h_sc, sizeof(Scalars))) != cudaSuccess) return bail(e, "cudaMallocHost");
    if ((e = cudaMallocHost((void **)&ctx->h_pr, sizeof(Params))) != cudaSuccess) return bail(e, "cudaMallocHost");
    cudaMemset(ctx->d_sc, 0, sizeof(Scalars));
    cudaMemset(ctx->d_pr, 0, sizeof(Params));
    memset(ctx->h_sc, 0, sizeof(Scalars));
    md_lj_new(ctx->sigma, ctx->eps, &ctx->r_cut, &ctx->u_cut);  // PotentialsDatabase::new()
    *out = ctx;
    return MD_OK;
}

void md_destroy(md_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    drop_graph(ctx);
    drop_chunk_cache(ctx);
    for (int k = 0; k < ctx->dist.n_ipc_opened; ++k) cudaIpcCloseMemHandle(ctx->dist.ipc_opened[k]);
    if (ctx->dist.comm) ncclCommDestroy(ctx->dist.comm);
    if (ctx->dist.h_cnt) cudaFreeHost(ctx->dist.h_cnt);
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    for (auto &b : ctx->owned)
        if (b.p) cudaFree(b.p);
    if (ctx->d_sc) cudaFree(ctx->d_sc);
    if (ctx->d_pr) cudaFree(ctx->d_pr);
    if (ctx->h_sc) cudaFreeHost(ctx->h_sc);
    if (ctx->h_pr) cudaFreeHost(ctx->h_pr);
    if (ctx->multi.h_tab) cudaFreeHost(ctx->multi.h_tab);
    if (ctx->multi.h_work) cudaFreeHost(ctx->multi.h_work);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int md_set_potential_lj(md_ctx *ctx, double sigma, double eps, double r_cut, double u_cut)
{
    TRY(check_ctx(ctx, false));
    if (!(sigma > 0.0) || !(r_cut > 0.0) || !std::isfinite(eps) || !std::isfinite(u_cut))
        return ctx->fail(MD_ERR_INVALID_ARGUMENT, "bad Lennard-Jones parameters");
    ctx->sigma = sigma; ctx->eps = eps; ctx->r_cut = r_cut; ctx->u_cut = u_cut;
    ctx->list_valid = false;
    ctx->force_valid = false;
    if (ctx->multi.on) ctx->multi.rc_max = multi_rc_max(ctx);  // pair (0, 0) of the type table: the list radius may change
    if (ctx->has_state) {
        TRY(pull_scalars(ctx));
        ctx->skin = choose_skin(ctx, ctx->h_sc->box);
    }
    fill_potential_params(ctx);
    return MD_OK;
}

int md_set_potential_pair(md_ctx *ctx, int32_t id0, int32_t id1, double sigma, double eps, double r_cut, double u_cut)
{
    TRY(check_ctx(ctx, false));
    if (id0 < 0 || id1 < 0 || id0 >= MULTI_MAX_TYPES || id1 >= MULTI_MAX_TYPES)
        return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_set_potential_pair: type ids must be in [0, %d)", MULTI_MAX_TYPES);
    if (id0 == 0 && id1 == 0) return md_set_potential_lj(ctx, sigma, eps, r_cut, u_cut);
    if (!(sigma > 0.0) || !(r_cut > 0.0) || !std::isfinite(eps) || !std::isfinite(u_cut))
        return ctx->fail(MD_ERR_INVALID_ARGUMENT, "bad Lennard-Jones parameters");
    ctx->multi.pairs[{std::min(id0, id1), std::max(id0, id1)}] = {sigma, eps, r_cut, u_cut};  // potential.rs:141-144
    ctx->list_valid = false;
    ctx->force_valid = false;
    if (ctx->multi.on) {
        ctx->multi.rc_max = multi_rc_max(ctx);
        if (ctx->has_state) {
            TRY(pull_scalars(ctx));
            ctx->skin = choose_skin(ctx, ctx->h_sc->box);
        }
        fill_potential_params(ctx);
    }
    return MD_OK;
}

int md_set_cross_type_mode(md_ctx *ctx, int32_t mode)
{
    TRY(check_ctx(ctx, false));
    if (mode != MD_CROSS_REFERENCE && mode != MD_CROSS_SYMMETRIC)
        return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_set_cross_type_mode: MD_CROSS_REFERENCE or MD_CROSS_SYMMETRIC");
    ctx->multi.mode = mode;
    ctx->force_valid = false;
    return MD_OK;
}

// (re)allocates the device-resident State of one GPU for n atoms
static int alloc_state(md_ctx *ctx, int64_t n)
{
    cudaStream_t st = ctx->stream;
    if (n != ctx->n || !ctx->has_state) {
        CK(cudaStreamSynchronize(st));
        drop_graph(ctx);
        free_arrays(ctx, &ctx->cur); free_arrays(ctx, &ctx->alt);
        dev_free(ctx, ctx->stage); dev_free(ctx, ctx->stage_i); dev_free(ctx, ctx->d_partials);
        dev_free(ctx, ctx->cell_of); dev_free(ctx, ctx->cell_sorted); dev_free(ctx, ctx->order);
        dev_free(ctx, ctx->nbr_cnt); dev_free(ctx, ctx->nbr);
        ctx->nbr = nullptr; ctx->nbr_alloc = 0;
        ctx->owned.erase(std::remove_if(ctx->owned.begin(), ctx->owned.end(), [](const DevBuf &b) { return !b.p; }),
                         ctx->owned.end());
        ctx->n = n;
        ctx->n_own = n;
        ctx->n_ghost = 0;
        ctx->npad = (int)((n + 64) / 64 * 64);  // at least one spare slot: the tile kernels' bulk copies end on even indices
        TRY(alloc_arrays(ctx, &ctx->cur, ctx->npad));
        TRY(alloc_arrays(ctx, &ctx->alt, ctx->npad));
        TRY(dev_alloc(ctx, &ctx->stage, 3 * (size_t)ctx->npad));
        TRY(dev_alloc(ctx, &ctx->stage_i, ctx->npad));
        TRY(choose_grids(ctx));
        ctx->partial_blocks = std::max(std::max(ctx->force_grid[0], ctx->force_grid[1]),
                                       std::max(ctx->force_grid[2], ctx->reduce_grid));
        ctx->partial_blocks = std::max(ctx->partial_blocks, ctx->loop_blocks_max);
        TRY(dev_alloc(ctx, &ctx->d_partials, (size_t)ctx->partial_blocks * NSUM));
        TRY(dev_alloc(ctx, &ctx->cell_of, ctx->npad));
        TRY(dev_alloc(ctx, &ctx->cell_sorted, ctx->npad));
        TRY(dev_alloc(ctx, &ctx->order, ctx->npad));
        TRY(dev_alloc(ctx, &ctx->nbr_cnt, ctx->npad));
        ctx->grid = Grid{};
    }
    return MD_OK;
}

// control words, skin, neighbour capacity and the macro parameters of a freshly written State (one GPU)
static int finish_new_state(md_ctx *ctx, int64_t n, const double box[3])
{
    cudaStream_t st = ctx->stream;
    Scalars &h = *ctx->h_sc;
    memset(&h, 0, sizeof h);
    h.box[0] = box[0]; h.box[1] = box[1]; h.box[2] = box[2];
    h.mu_pending = 1.0; h.lambda = 1.0; h.mu = 1.0; h.inv_scale = 1.0;
    h.lambda_last = 1.0; h.mu_last = 1.0;
    CK(cudaMemcpyAsync(ctx->d_sc, ctx->h_sc, sizeof(Scalars), cudaMemcpyHostToDevice, st));
    drop_graph(ctx);
    ctx->tile_disabled = false;
    ctx->tile_valid = false;
    ctx->cap16 = 0;
    ctx->skin = choose_skin(ctx, box);
    fill_potential_params(ctx);
    // neighbour capacity from density
    {
        double volume = box[0] * box[1] * box[2];
        double r_list = list_cut(ctx) + ctx->skin;
        double expect = (double)n / volume * 4.18879020478639 * r_list * r_list * r_list;
        int cap = ctx->cfg.max_neighbours > 0 ? ctx->cfg.max_neighbours : (int)(expect * 1.5) + 16;
        cap = std::min<int64_t>((cap + 7) / 8 * 8, std::max<int64_t>(8, (n - 1 + 7) / 8 * 8));
        TRY(ensure_nbr_capacity(ctx, cap));
    }
    TRY(push_params(ctx));
    // macro parameters of the uploaded state: two passes so the thermal sum is shifted by the true COM velocity
    launch_reduce(ctx);
    k_set_shift_to_vcom<<<1, 1, 0, st>>>(ctx->d_sc);
    launch_reduce(ctx);
    ctx->stats.kernel_launches += 1;
    ctx->sums_c = ctx->prm.half_dt_m;
    ctx->sums_nh = true;
    CK(cudaGetLastError());
    return MD_OK;
}

int md_upload_state(md_ctx *ctx, int64_t n, const double *pos, const double *vel, const double *force,
                    const double *potential, const double *virial, double mass, const double box[3])
{
    TRY(check_ctx(ctx, false));
    if (n <= 0 || n > (int64_t)1 << 30 || !pos || !vel || !box)
        return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_upload_state: need n in [1, 2^30], pos, vel, box");
    if (!(mass > 0.0) || !(box[0] > 0.0) || !(box[1] > 0.0) || !(box[2] > 0.0))
        return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_upload_state: mass and box must be positive");
    if (ctx->dist.on) return dist_upload(ctx, n, pos, vel, force, potential, virial, mass, box);
    cudaStream_t st = ctx->stream;
    ctx->multi.on = false;
    TRY(alloc_state(ctx, n));
    ctx->mass = mass;
    ctx->has_state = true;
    ctx->list_valid = false;
    ctx->force_valid = force != nullptr;
    const int ni = (int)n;
    const int nb = blocks_for(n, 256);
    const size_t b3 = 3 * (size_t)n * sizeof(double), b1 = (size_t)n * sizeof(double);
    CK(cudaMemcpyAsync(ctx->stage, pos, b3, cudaMemcpyHostToDevice, st));
    k_deinterleave3<<<nb, 256, 0, st>>>(ni, ctx->stage, ctx->cur.x, ctx->cur.y, ctx->cur.z);
    CK(cudaMemcpyAsync(ctx->stage, vel, b3, cudaMemcpyHostToDevice, st));
    k_deinterleave3<<<nb, 256, 0, st>>>(ni, ctx->stage, ctx->cur.vx, ctx->cur.vy, ctx->cur.vz);
    if (force) {
        CK(cudaMemcpyAsync(ctx->stage, force, b3, cudaMemcpyHostToDevice, st));
        k_deinterleave3<<<nb, 256, 0, st>>>(ni, ctx->stage, ctx->cur.fx, ctx->cur.fy, ctx->cur.fz);
    } else {
        CK(cudaMemsetAsync(ctx->cur.fx, 0, b1, st));
        CK(cudaMemsetAsync(ctx->cur.fy, 0, b1, st));
        CK(cudaMemsetAsync(ctx->cur.fz, 0, b1, st));
    }
    if (potential) CK(cudaMemcpyAsync(ctx->cur.u, potential, b1, cudaMemcpyHostToDevice, st));
    else CK(cudaMemsetAsync(ctx->cur.u, 0, b1, st));
    if (virial) CK(cudaMemcpyAsync(ctx->cur.w, virial, b1, cudaMemcpyHostToDevice, st));
    else CK(cudaMemsetAsync(ctx->cur.w, 0, b1, st));
    k_iota<<<nb, 256, 0, st>>>(ni, ctx->cur.id);
    ctx->stats.kernel_launches += 3 + (force ? 1 : 0);

    TRY(finish_new_state(ctx, n, box));
    CK(cudaStreamSynchronize(st));  // host buffers are only borrowed for the duration of the call
    return MD_OK;
}


int md_upload_state_typed(md_ctx *ctx, int64_t n, const double *pos, const double *vel, const double *force,
                          const double *potential, const double *virial, int32_t n_types, const int64_t *type_counts,
                          const double *type_mass, const double box[3])
{
    TRY(check_ctx(ctx, false));
    if (n_types < 1 || n_types > MULTI_MAX_TYPES || !type_counts || !type_mass)
        return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_upload_state_typed: 1..%d particle types with counts and masses", MULTI_MAX_TYPES);
    int64_t total = 0;
    for (int t = 0; t < n_types; ++t) {
        // an empty type makes the reference panic (particle_type[0], integrator.rs:29)
        if (type_counts[t] <= 0 || !(type_mass[t] > 0.0))
            return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_upload_state_typed: type %d needs at least one atom and a positive mass", t);
        total += type_counts[t];
    }
    if (total != n) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_upload_state_typed: the type counts do not add up to n");
    if (ctx->dist.on) return ctx->fail(MD_ERR_UNSUPPORTED, "several particle types on a decomposed context");
    if (n_types == 1) return md_upload_state(ctx, n, pos, vel, force, potential, virial, type_mass[0], box);
    // the single-type upload lays out the planes; then the type table takes over (list radius, skin, capacity)
    TRY(md_upload_state(ctx, n, pos, vel, force, potential, virial, type_mass[0], box));
    auto &m = ctx->multi;
    m.T = n_types;
    m.start[0] = 0;
    for (int t = 0; t < n_types; ++t) {
        m.start[t + 1] = m.start[t] + type_counts[t];
        m.mass[t] = type_mass[t];
    }
    if (!m.d_tab) {
        TRY(dev_alloc(ctx, &m.d_tab, 1));
        TRY(dev_alloc(ctx, &m.d_work, 1));
        CK(cudaHostAlloc((void **)&m.h_tab, sizeof(MultiTable), cudaHostAllocDefault));
        CK(cudaHostAlloc((void **)&m.h_work, sizeof(MultiWork), cudaHostAllocDefault));
    }
    const int nblocks = std::max(1, std::min(blocks_for(n, MULTI_BLOCK), 296));
    if (nblocks > m.nblocks) {
        dev_free(ctx, m.d_partials);
        TRY(dev_alloc(ctx, &m.d_partials, (size_t)MULTI_MAX_TYPES * nblocks * MULTI_NA));
    }
    m.nblocks = nblocks;
    m.on = true;
    m.disp = 0.0;
    m.rc_max = multi_rc_max(ctx);
    ctx->skin = choose_skin(ctx, box);
    fill_potential_params(ctx);
    {
        const double volume = box[0] * box[1] * box[2], r_list = m.rc_max + ctx->skin;
        const double expect = (double)n / volume * 4.18879020478639 * r_list * r_list * r_list;
        int cap = ctx->cfg.max_neighbours > 0 ? ctx->cfg.max_neighbours : (int)(expect * 1.5) + 16;
        cap = (int)std::min<int64_t>((cap + 7) / 8 * 8, std::max<int64_t>(8, (n - 1 + 7) / 8 * 8));
        TRY(ensure_nbr_capacity(ctx, cap));
    }
    ctx->tile_disabled = true;  // (the brick kernels are single-type)
    ctx->list_valid = false;
    CK(cudaStreamSynchronize(ctx->stream));
    return MD_OK;
}

int md_initialize_lattice(md_ctx *ctx, int cell_type, const int32_t size[3], const double start[3], double unit_cell,
                          double mass, double temperature, uint64_t seed)
{
    TRY(check_ctx(ctx, false));
    if (!size || size[0] <= 0 || size[1] <= 0 || size[2] <= 0 || !(unit_cell > 0.0) || !(mass > 0.0) ||
        !(temperature >= 0.0) || (cell_type != MD_CELL_UNIFORM && cell_type != MD_CELL_FCC))
        return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_initialize_lattice: bad arguments");
    if (ctx->dist.on)
        return ctx->fail(MD_ERR_UNSUPPORTED, "md_initialize_lattice on a decomposed context: initialise on one GPU or upload");
    const int64_t cells = (int64_t)size[0] * size[1] * size[2];
    const int64_t n = cells * (cell_type == MD_CELL_FCC ? 4 : 1);
    if (n > (int64_t)1 << 30) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_initialize_lattice: more than 2^30 atoms");
    const double box[3] = {unit_cell * (double)size[0], unit_cell * (double)size[1], unit_cell * (double)size[2]};
    const double s0[3] = {start ? start[0] : 0.0, start ? start[1] : 0.0, start ? start[2] : 0.0};
    cudaStream_t st = ctx->stream;
    ctx->multi.on = false;
    TRY(alloc_state(ctx, n));
    ctx->mass = mass;
    ctx->has_state = true;
    ctx->list_valid = false;
    ctx->force_valid = false;
    const int ni = (int)n, nb = blocks_for(n, 256);
    k_init_lattice<<<blocks_for(cells, 256), 256, 0, st>>>((int)cells, cell_type == MD_CELL_FCC ? 1 : 0, size[0], size[1],
                                                         size[2], s0[0], s0[1], s0[2], unit_cell, ctx->cur);
    // velocity.rs:8-11: sigma = sqrt(K_B * (T * 0.01) / mass)
    const double sigma = std::sqrt(K_B * (temperature * 0.01) / mass);
    k_init_velocities<<<nb, 256, 0, st>>>(ni, sigma, (unsigned long long)seed, ctx->cur);
    k_iota<<<nb, 256, 0, st>>>(ni, ctx->cur.id);
    ctx->stats.kernel_launches += 3;
    TRY(finish_new_state(ctx, n, box));
    CK(cudaStreamSynchronize(st));
    return MD_OK;
}

int md_download_state(md_ctx *ctx, double *pos, double *vel, double *force, double *potential, double *virial,
                      double box[3])
{
    TRY(check_ctx(ctx, true));
    if (ctx->dist.on)
        return ctx->fail(MD_ERR_UNSUPPORTED, "md_download_state on a decomposed state: use md_download_local per rank");
    cudaStream_t st = ctx->stream;
    const int n = (int)ctx->n;
    const int nb = blocks_for(n, 256);
    const size_t b3 = 3 * (size_t)n * sizeof(double), b1 = (size_t)n * sizeof(double);
    const Arrays &a = ctx->cur;
    if (pos) {
        k_interleave3_unsort<<<nb, 256, 0, st>>>(n, a.x, a.y, a.z, a.id, ctx->stage);
        CK(cudaMemcpyAsync(pos, ctx->stage, b3, cudaMemcpyDeviceToHost, st));
    }
    if (vel) {
        k_interleave3_unsort<<<nb, 256, 0, st>>>(n, a.vx, a.vy, a.vz, a.id, ctx->stage);
        CK(cudaMemcpyAsync(vel, ctx->stage, b3, cudaMemcpyDeviceToHost, st));
    }
    if (force) {
        k_interleave3_unsort<<<nb, 256, 0, st>>>(n, a.fx, a.fy, a.fz, a.id, ctx->stage);
        CK(cudaMemcpyAsync(force, ctx->stage, b3, cudaMemcpyDeviceToHost, st));
    }
    if (potential) {
        k_unsort1<<<nb, 256, 0, st>>>(n, a.u, a.id, ctx->stage);
        CK(cudaMemcpyAsync(potential, ctx->stage, b1, cudaMemcpyDeviceToHost, st));
    }
    if (virial) {
        k_unsort1<<<nb, 256, 0, st>>>(n, a.w, a.id, ctx->stage);
        CK(cudaMemcpyAsync(virial, ctx->stage, b1, cudaMemcpyDeviceToHost, st));
    }
    ctx->stats.kernel_launches += (pos != nullptr) + (vel != nullptr) + (force != nullptr) + (potential != nullptr) +
                                  (virial != nullptr);
    CK(cudaGetLastError());
    TRY(pull_scalars(ctx));
    if (box) {
        box[0] = ctx->h_sc->box[0]; box[1] = ctx->h_sc->box[1]; box[2] = ctx->h_sc->box[2];
    }
    return MD_OK;
}

int md_update_force(md_ctx *ctx)
{
    TRY(check_ctx(ctx, true));
    if (ctx->multi.on) return multi_update_force(ctx);
    fill_potential_params(ctx);
    TRY(push_params(ctx));
    if (ctx->dist.on) {
        if (!ctx->list_valid) {
            TRY(dist_rebuild(ctx));
        } else {
            TRY(pull_scalars(ctx));
            if (ctx->h_sc->need_rebuild) TRY(dist_rebuild(ctx));
            else TRY(dist_halo_exchange(ctx));
        }
        TRY(dist_launch_force(ctx, false));
        ctx->sums_c = ctx->prm.half_dt_m;
        ctx->sums_nh = true;
        ctx->force_valid = true;
        CK(cudaStreamSynchronize(ctx->stream));
        return MD_OK;
    }
    if (!ctx->list_valid) {
        TRY(rebuild_lists(ctx));
    } else {
        TRY(pull_scalars(ctx));
        if (ctx->h_sc->need_rebuild) TRY(rebuild_lists(ctx));
    }
    launch_force(ctx, false, 0ull);
    ctx->stats.kernel_launches += 1;
    CK(cudaGetLastError());
    ctx->sums_c = ctx->prm.half_dt_m;
    ctx->sums_nh = true;
    ctx->force_valid = true;
    CK(cudaStreamSynchronize(ctx->stream));
    return MD_OK;
}

int md_step(md_ctx *ctx, int64_t n_steps, double dt, md_thermostat *th, md_barostat *ba)
{
    TRY(check_ctx(ctx, true));
    if (n_steps < 0 || !std::isfinite(dt)) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_step: bad n_steps/dt");
    if (th && th->kind != MD_THERMOSTAT_NONE && th->kind != MD_THERMOSTAT_BERENDSEN &&
        th->kind != MD_THERMOSTAT_NOSE_HOOVER)
        return ctx->fail(MD_ERR_UNSUPPORTED, "Thermostat::Custom is todo!() in the reference");
    if (ba && ba->kind != MD_BAROSTAT_NONE && ba->kind != MD_BAROSTAT_BERENDSEN)
        return ctx->fail(MD_ERR_UNSUPPORTED, "Barostat::Custom is todo!() in the reference");
    if (n_steps == 0) return MD_OK;
    if (ctx->multi.on) return multi_step(ctx, n_steps, dt, th, ba);
    cudaStream_t st = ctx->stream;

    fill_potential_params(ctx);
    Params &p = ctx->prm;
    p.dt = dt;
    p.half_dt_m = dt / (2.0 * ctx->mass);  // integrator.rs:30
    p.th_kind = th ? th->kind : 0;
    p.th_tau = th ? th->tau : 1.0;
    p.th_target = th ? th->target : 0.0;
    p.ba_kind = ba ? ba->kind : 0;
    p.ba_beta = ba ? ba->beta : 0.0;
    p.ba_tau = ba ? ba->tau : 1.0;
    p.ba_target = ba ? ba->target : 0.0;
    TRY(push_params(ctx));
    if (ctx->graph_hc != p.half_dt_m) {  // dt/(2m) is a launch constant of the captured force kernel
        drop_graph(ctx);
        ctx->graph_hc = p.half_dt_m;
    }
    // The displacement bound of the first drift needs max|v + F c|² for THIS c.
    if (ctx->sums_c != p.half_dt_m || (p.th_kind == MD_THERMOSTAT_NOSE_HOOVER && !ctx->sums_nh)) {
        if (ctx->dist.on) TRY(dist_launch_reduce(ctx));
        else launch_reduce(ctx);
        ctx->sums_c = p.half_dt_m;
        ctx->sums_nh = true;
    }
    ctx->sums_nh = p.th_kind == MD_THERMOSTAT_NOSE_HOOVER;  // what the step kernels of this batch will leave behind
    k_prepare<<<1, 1, 0, st>>>(ctx->d_sc, ctx->d_pr, (long long)n_steps, th ? th->psi : 0.0);
    ctx->stats.kernel_launches += 1;
    CK(cudaGetLastError());

    // The step loop.  The device knows whether it may step — list still valid, no error, steps left — and every step kernel
    // checks that itself (a launch that may not step is a no-op).  So the host does not ask first: while its picture of the
    // device is stale it enqueues steps speculatively and looks afterwards; only a rebuild step, which the host runs by hand,
    // needs the picture to be fresh.  One host round trip per md_step call or list epoch instead of three.
    int rc = MD_OK;
    bool fresh = false;               // h_sc mirrors the device's controls
    long long bound = n_steps;        // upper bound of the steps left on the device
    long long seen_done = 0;          // steps_done at the last look (k_prepare reset it)
    unsigned long long loop_ns0[4] = {0, 0, 0, 0}, loop_steps0 = 0;  // the persistent loop's phase clocks before this batch
    bool have_loop0 = false;
    auto look = [&]() -> int {
        TRY(pull_scalars(ctx));
        fresh = true;
        const long long ran = ctx->h_sc->steps_done - seen_done;
        seen_done = ctx->h_sc->steps_done;
        ctx->stats.steps += ran;
        return device_error(ctx);
    };
    if (ctx->timing || !ctx->list_valid) {
        rc = look();
        if (rc == MD_OK) {
            for (int k = 0; k < 4; ++k) loop_ns0[k] = ctx->h_sc->loop_ns[k];
            loop_steps0 = ctx->h_sc->loop_steps;
            have_loop0 = true;
        }
    }
    while (rc == MD_OK) {
        const bool use_loop = loop_wanted(ctx) && ctx->loop_blocks_max > 0;
        // (the persistent loop times its phases itself; the two-kernel step is timed with events around host-stepped launches)
        const bool host_stepped = ctx->cfg.loop_mode == MD_LOOP_HOST || (ctx->timing && !use_loop);
        const bool by_hand = host_stepped && !use_loop && !ctx->dist.on;
        if (fresh) {
            bound = ctx->h_sc->steps_left;
            if (bound <= 0) break;
            const bool rebuild = !ctx->list_valid || ctx->h_sc->need_rebuild;
            if (rebuild || by_hand) {
                // one step by hand: drift, (rebuild at the drifted positions,) forces
                if (ctx->timing) CK(cudaEventRecord(ctx->ev[0], st));
                launch_kick_drift(ctx);
                if (ctx->timing) CK(cudaEventRecord(ctx->ev[1], st));
                if (rebuild && (rc = ctx->dist.on ? dist_rebuild(ctx) : rebuild_lists(ctx)) != MD_OK) break;
                if (ctx->timing) CK(cudaEventRecord(ctx->ev[2], st));
                if (ctx->dist.on) {
                    if ((rc = dist_launch_force(ctx, true)) != MD_OK) break;
                } else {
                    launch_force(ctx, true);
                }
                if (ctx->timing) CK(cudaEventRecord(ctx->ev[3], st));
                ctx->stats.kernel_launches += 2;
                CK(cudaGetLastError());
                if (ctx->timing) {
                    CK(cudaEventSynchronize(ctx->ev[3]));
                    float a = 0.f, b = 0.f, c = 0.f;
                    CK(cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]));
                    CK(cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]));
                    CK(cudaEventElapsedTime(&c, ctx->ev[2], ctx->ev[3]));
                    if (rebuild) { ctx->t_ms[2] += b; ctx->t_cnt[2] += 1; }
                    else {
                        ctx->t_ms[0] += a; ctx->t_cnt[0] += 1;
                        ctx->t_ms[1] += c; ctx->t_cnt[1] += 1;
                    }
                }
                fresh = false;
                bound -= 1;
                if (bound <= 0 || by_hand) rc = look();
                continue;
            }
        }
        if (by_hand) {  // (one host round trip per step is what this mode is for)
            rc = look();
            continue;
        }
        // list valid as far as the host knows: enqueue, then look
        if (use_loop) {
            // dilute systems: the persistent step loop runs until the device asks for a rebuild or the batch is done (on
            // several GPUs every rank launches its loop; the loops leave at the same step: globally agreed controls)
            if ((rc = launch_loop(ctx, host_stepped ? 1 : bound)) != MD_OK) break;
            ctx->stats.kernel_launches += 1;
            rc = look();
        } else if (ctx->dist.on) {
            const long long before = seen_done;
            if ((rc = dist_enqueue_chunks(ctx, bound)) != MD_OK) break;
            if ((rc = look()) != MD_OK) break;
            ctx->stats.kernel_launches += (seen_done - before) * (ctx->dist.p2p ? 2 : 7);  // kick_drift, force, pack x2, unpack x2, finalize
        } else {
            cudaGraphExec_t chunk_exec = nullptr;
            if ((rc = get_chunk_graph(ctx, &chunk_exec)) != MD_OK) break;
            // look ahead as far as the list is expected to last (the previous epoch's length), at most 8 chunks: steps
            // enqueued past a rebuild request are no-ops, but each still costs a launch
            const long long since = ctx->stats.steps - ctx->epoch_start_step;
            const long long expect = std::max<long long>(STEP_CHUNK, ctx->last_epoch_len - since);
            const int chunks = (int)std::min<long long>(8, (std::min<long long>(bound, expect) + STEP_CHUNK - 1) / STEP_CHUNK);
            for (int c = 0; c < chunks; ++c) CK(cudaGraphLaunch(chunk_exec, st));
            ctx->stats.graph_launches += chunks;
            const long long before = seen_done;
            if ((rc = look()) != MD_OK) break;
            ctx->stats.kernel_launches += 2 * (seen_done - before);
        }
    }
    if (rc == MD_OK && p.ba_kind == MD_BAROSTAT_BERENDSEN) {
        rc = flush_pending_scale(ctx);
        fresh = false;
    }
    if (rc == MD_OK && !fresh) rc = look();
    if (rc != MD_OK) {
        // The device may be mid-batch: velocity planes holding u = v + F c, a pending coordinate scale, stale F/U/W.  Nothing
        // the host could download would be a State of the reference's step sequence: the caller has to upload again.
        ctx->has_state = false;
        ctx->list_valid = false;
        ctx->force_valid = false;
        cudaStreamSynchronize(st);
        return rc;
    }
    if (ctx->timing && have_loop0) {
        const long long ls = (long long)(ctx->h_sc->loop_steps - loop_steps0);
        if (ls > 0) {
            ctx->t_ms[0] += (double)(ctx->h_sc->loop_ns[0] - loop_ns0[0]) * 1e-6; ctx->t_cnt[0] += ls;
            ctx->t_ms[1] += (double)((ctx->h_sc->loop_ns[2] - loop_ns0[2]) + (ctx->h_sc->loop_ns[3] - loop_ns0[3])) * 1e-6;
            ctx->t_cnt[1] += ls;
            ctx->t_ms[3] += (double)(ctx->h_sc->loop_ns[1] - loop_ns0[1]) * 1e-6; ctx->t_cnt[3] += ls;
        }
    }
    if (th) {
        th->lambda = ctx->h_sc->lambda_last;
        if (th->kind == MD_THERMOSTAT_NOSE_HOOVER) th->psi = ctx->h_sc->psi;
    }
    if (ba) ba->myu = ctx->h_sc->mu_last;
    ctx->force_valid = true;
    return MD_OK;
}

int md_time_kernels(md_ctx *ctx, int64_t n_steps, double dt, md_thermostat *th, md_barostat *ba, double ms[4],
                    int64_t launches[4])
{
    TRY(check_ctx(ctx, true));
    if (!ms || !launches) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_time_kernels: NULL output");
    for (auto &e : ctx->ev)
        if (!e) CK(cudaEventCreate(&e));
    for (int k = 0; k < 4; ++k) { ctx->t_ms[k] = 0.0; ctx->t_cnt[k] = 0; }
    ctx->timing = true;
    int rc = md_step(ctx, n_steps, dt, th, ba);
    ctx->timing = false;
    for (int k = 0; k < 4; ++k) { ms[k] = ctx->t_ms[k]; launches[k] = ctx->t_cnt[k]; }
    return rc;
}

int md_comm_unique_id(uint8_t id[MD_UNIQUE_ID_BYTES])
{
    static_assert(sizeof(ncclUniqueId) == MD_UNIQUE_ID_BYTES, "ncclUniqueId size");
    if (!id) return MD_ERR_INVALID_ARGUMENT;
    if (const char *e = nccl_api::load()) {
        g_create_error = e;
        return MD_ERR_NCCL;
    }
    ncclUniqueId u;
    if (ncclGetUniqueId(&u) != ncclSuccess) {
        g_create_error = "ncclGetUniqueId failed";
        return MD_ERR_NCCL;
    }
    memcpy(id, &u, sizeof u);
    return MD_OK;
}

int md_comm_init(md_ctx *ctx, int rank, int nranks, const uint8_t id[MD_UNIQUE_ID_BYTES])
{
    TRY(check_ctx(ctx, false));
    if (nranks < 1 || rank < 0 || rank >= nranks || !id) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_comm_init: bad rank/nranks/id");
    if (ctx->has_state) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_comm_init must precede md_upload_state");
    if (nranks == 1) return MD_OK;  // a single slab is the single-GPU path
    if (const char *e = nccl_api::load()) return ctx->fail(MD_ERR_NCCL, "%s", e);
    auto &d = ctx->dist;
    ncclUniqueId u;
    memcpy(&u, id, sizeof u);
    NCK(ncclCommInitRank(&d.comm, nranks, u, rank));
    d.on = true;
    d.rank = rank;
    d.nranks = nranks;
    d.left = (rank + nranks - 1) % nranks;
    d.right = (rank + 1) % nranks;
    TRY(dev_alloc(ctx, &d.d_cnt, 8));
    CK(cudaMallocHost((void **)&d.h_cnt, 8 * sizeof(int)));
    return MD_OK;
}

int md_local_count(md_ctx *ctx, int64_t *n_owned, int64_t *n_ghost)
{
    TRY(check_ctx(ctx, true));
    if (n_owned) *n_owned = ctx->n_own;
    if (n_ghost) *n_ghost = ctx->n_ghost;
    return MD_OK;
}

int md_download_local(md_ctx *ctx, int64_t *ids, double *pos, double *vel, double *force, double *potential,
                      double *virial, double box[3])
{
    TRY(check_ctx(ctx, true));
    cudaStream_t st = ctx->stream;
    const int n = (int)ctx->n_own;
    const int nb = std::max(1, blocks_for(n, 256));
    const size_t b3 = 3 * (size_t)n * sizeof(double), b1 = (size_t)n * sizeof(double);
    const Arrays &a = ctx->cur;
    if (ctx->h_sc->mu_pending != 1.0) { /* pending scaling is flushed at the end of md_step */ }
    if (pos && n) {
        k_interleave3<<<nb, 256, 0, st>>>(n, a.x, a.y, a.z, ctx->stage);
        CK(cudaMemcpyAsync(pos, ctx->stage, b3, cudaMemcpyDeviceToHost, st));
    }
    if (vel && n) {
        k_interleave3<<<nb, 256, 0, st>>>(n, a.vx, a.vy, a.vz, ctx->stage);
        CK(cudaMemcpyAsync(vel, ctx->stage, b3, cudaMemcpyDeviceToHost, st));
    }
    if (force && n) {
        k_interleave3<<<nb, 256, 0, st>>>(n, a.fx, a.fy, a.fz, ctx->stage);
        CK(cudaMemcpyAsync(force, ctx->stage, b3, cudaMemcpyDeviceToHost, st));
    }
    if (potential && n) CK(cudaMemcpyAsync(potential, a.u, b1, cudaMemcpyDeviceToHost, st));
    if (virial && n) CK(cudaMemcpyAsync(virial, a.w, b1, cudaMemcpyDeviceToHost, st));
    if (ids && n) {
        std::vector<int> tmp((size_t)n);
        CK(cudaMemcpyAsync(tmp.data(), a.id, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        for (int i = 0; i < n; ++i) ids[i] = tmp[(size_t)i];
    }
    CK(cudaGetLastError());
    TRY(pull_scalars(ctx));
    if (box) {
        box[0] = ctx->h_sc->box[0]; box[1] = ctx->h_sc->box[1]; box[2] = ctx->h_sc->box[2];
    }
    return MD_OK;
}

// Host-only description of the slab a rank owns (fractional cut along x) and a capacity hint; no GPU involved.
int md_plan_decomposition(int64_t n, const double box[3], double r_list, int nranks, int rank, double *x_lo,
                          double *x_hi, int *left, int *right, int64_t *capacity_hint)
{
    if (n <= 0 || !box || nranks < 1 || rank < 0 || rank >= nranks || !(r_list > 0.0)) return MD_ERR_INVALID_ARGUMENT;
    if (nranks > 1 && box[0] / nranks < 2.1 * r_list) return MD_ERR_DECOMPOSITION;
    if (x_lo) *x_lo = box[0] * ((double)rank / (double)nranks);
    if (x_hi) *x_hi = box[0] * ((double)(rank + 1) / (double)nranks);
    if (left) *left = (rank + nranks - 1) % nranks;
    if (right) *right = (rank + 1) % nranks;
    if (capacity_hint) {
        double own = (double)n / nranks, ghosts = nranks > 1 ? 2.0 * (double)n * r_list / box[0] : 0.0;
        *capacity_hint = (int64_t)(1.5 * own + 3.0 * ghosts) + 4096;
    }
    return MD_OK;
}

int md_macro(md_ctx *ctx, md_macro_out *out)
{
    TRY(check_ctx(ctx, true));
    if (!out) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_macro: out is NULL");
    if (ctx->multi.on)
        return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_macro: the State has several particle types — the reference's macro "
                                                  "parameters are per type: md_macro_type");
    TRY(pull_scalars(ctx));
    const Scalars &h = *ctx->h_sc;
    out->kinetic_energy = h.kinetic;
    out->thermal_energy = h.thermal;
    out->potential_energy = h.potential;
    out->temperature = h.temperature;
    out->pressure = h.pressure;
    for (int d = 0; d < 3; ++d) {
        out->vcom[d] = h.vcom[d];
        out->momentum[d] = h.sum_mv[d];
        out->box[d] = h.box[d];
    }
    out->lambda = h.lambda_last;
    out->myu = h.mu_last;
    out->n = ctx->n;
    return MD_OK;
}

int md_macro_type(md_ctx *ctx, int32_t type_id, md_macro_out *out)
{
    TRY(check_ctx(ctx, true));
    if (!out) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_macro_type: out is NULL");
    if (!ctx->multi.on) {
        if (type_id != 0) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_macro_type: the State has one particle type");
        return md_macro(ctx, out);
    }
    auto &m = ctx->multi;
    if (type_id < 0 || type_id >= m.T) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_macro_type: no such particle type");
    TRY(multi_push_table(ctx, ctx->prm.dt));
    TRY(multi_reduce(ctx, 0));
    CK(cudaMemcpyAsync(m.h_work, m.d_work, sizeof(MultiWork), cudaMemcpyDeviceToHost, ctx->stream));
    TRY(pull_scalars(ctx));
    const MultiTypeSums &ts = m.h_work->type[type_id];
    out->kinetic_energy = ts.kinetic;
    out->thermal_energy = ts.thermal;
    out->potential_energy = ts.potential;
    out->temperature = ts.temperature;
    out->pressure = ts.pressure;
    for (int d = 0; d < 3; ++d) {
        out->vcom[d] = ts.vcom[d];
        out->momentum[d] = ts.a[d] * m.mass[type_id];  // get_momentum_of_system mod.rs:28-34
        out->box[d] = ctx->h_sc->box[d];
    }
    out->lambda = ctx->h_sc->lambda_last;
    out->myu = ctx->h_sc->mu_last;
    out->n = m.start[type_id + 1] - m.start[type_id];
    return MD_OK;
}

int md_update_force_host(md_ctx *ctx, int64_t n, const double *pos, double mass, const double box[3], double *force,
                         double *potential, double *virial)
{
    TRY(check_ctx(ctx, false));
    if (n <= 0 || !pos) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_update_force_host: bad arguments");
    // velocities do not enter update_force; reuse pos as a stand-in so no extra host buffer is needed
    TRY(md_upload_state(ctx, n, pos, pos, nullptr, nullptr, nullptr, mass, box));
    TRY(md_update_force(ctx));
    return md_download_state(ctx, nullptr, nullptr, force, potential, virial, nullptr);
}

int md_calculate_host(md_ctx *ctx, int64_t n, double *pos, double *vel, double *force, double *potential,
                      double *virial, double mass, double box[3], double dt, md_thermostat *th, md_barostat *ba)
{
    TRY(check_ctx(ctx, false));
    if (!force) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_calculate_host: force is required (the step starts "
                                                          "with a half-kick from State.force)");
    // Particle.potential of the incoming State does not enter the step (update_force zeroes it, potential.rs:160-166) and
    // Particle.temp (the virial) only through barostat.calculate_myu's get_pressure (barostat.rs:23-29): what nothing
    // reads is not uploaded, only written back
    const bool need_virial = ba && ba->kind != MD_BAROSTAT_NONE;
    TRY(md_upload_state(ctx, n, pos, vel, force, nullptr, need_virial ? virial : nullptr, mass, box));
    TRY(md_step(ctx, 1, dt, th, ba));
    return md_download_state(ctx, pos, vel, force, potential, virial, box);
}

int md_download_cells(md_ctx *ctx, int32_t *cell_of_atom, int32_t dims[3])
{
    TRY(check_ctx(ctx, true));
    if (!ctx->list_valid) return ctx->fail(MD_ERR_NO_STATE, "no neighbour list built yet");
    const int n = (int)ctx->n;
    if (cell_of_atom) {
        k_unsort1i<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(n, ctx->cell_sorted, ctx->cur.id, ctx->stage_i);
        CK(cudaMemcpyAsync(cell_of_atom, ctx->stage_i, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (ctx->grid.brick)  // brick-major column numbering → the canonical (cx*ny + cy)*nz + cz
            for (int i = 0; i < n; ++i) {
                int cx, cy, cz;
                cell_decode(ctx->grid, cell_of_atom[i], cx, cy, cz);
                cell_of_atom[i] = (cx * ctx->grid.nc[1] + cy) * ctx->grid.nc[2] + cz;
            }
    }
    if (dims) {
        dims[0] = ctx->grid.nc[0]; dims[1] = ctx->grid.nc[1]; dims[2] = ctx->grid.nc[2];
    }
    return MD_OK;
}

static int fetch_lists(md_ctx *ctx, std::vector<int> &cnt, std::vector<int> &id, std::vector<int> *nbr)
{
    if (!ctx->list_valid) return ctx->fail(MD_ERR_NO_STATE, "no neighbour list built yet");
    const size_t n = (size_t)ctx->n;
    cnt.resize(n);
    id.resize(n);
    CK(cudaMemcpyAsync(cnt.data(), ctx->nbr_cnt, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaMemcpyAsync(id.data(), ctx->cur.id, sizeof(int) * n, cudaMemcpyDeviceToHost, ctx->stream));
    if (nbr && ctx->tile_valid) {
        // brick-local 16-bit lists → sorted indices in the k-major table the other paths use (introspection only)
        if ((size_t)ctx->cap16 * (size_t)ctx->npad > ctx->nbr_alloc || ctx->grid.cap < ctx->cap16)
            TRY(ensure_nbr_capacity(ctx, ctx->cap16));
        k_tile_expand<<<ctx->nbricks, TILE_BLOCK, (size_t)ctx->sh_cap * sizeof(int), ctx->stream>>>(
            ctx->grid, ctx->cell_start, ctx->d_sc, ctx->nbrT, ctx->cap16, ctx->nbr_cnt, ctx->nbr, ctx->npad);
        CK(cudaGetLastError());
    }
    if (nbr) {
        nbr->resize((size_t)ctx->grid.cap * ctx->npad);
        CK(cudaMemcpyAsync(nbr->data(), ctx->nbr, sizeof(int) * nbr->size(), cudaMemcpyDeviceToHost, ctx->stream));
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return MD_OK;
}

int md_neighbour_counts(md_ctx *ctx, int64_t *counts)
{
    TRY(check_ctx(ctx, true));
    if (!counts) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "counts is NULL");
    std::vector<int> cnt, id;
    TRY(fetch_lists(ctx, cnt, id, nullptr));
    for (size_t p = 0; p < cnt.size(); ++p) counts[id[p]] = cnt[p];
    return MD_OK;
}

int md_neighbour_lists(md_ctx *ctx, const int64_t *offsets, int64_t *partners)
{
    TRY(check_ctx(ctx, true));
    if (!offsets || !partners) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "offsets/partners is NULL");
    std::vector<int> cnt, id, nbr;
    TRY(fetch_lists(ctx, cnt, id, &nbr));
    for (size_t p = 0; p < cnt.size(); ++p) {
        int64_t *dst = partners + offsets[id[p]];
        for (int k = 0; k < cnt[p]; ++k) dst[k] = id[nbr[(size_t)k * ctx->npad + p]];
        std::sort(dst, dst + cnt[p]);
    }
    return MD_OK;
}

int md_get_stats(md_ctx *ctx, md_stats *out)
{
    if (!ctx || !out) return MD_ERR_INVALID_ARGUMENT;
    *out = ctx->stats;
    out->cells[0] = ctx->grid.nc[0]; out->cells[1] = ctx->grid.nc[1]; out->cells[2] = ctx->grid.nc[2];
    out->nbr_capacity = ctx->grid.cap;
    out->skin = ctx->skin;
    out->n_owned = ctx->n_own;
    out->n_ghost = ctx->n_ghost;
    out->migrated = ctx->dist.migrated;
    out->wait_halo_ms = (double)ctx->h_sc->wait_halo_ns * 1e-6;  // as of the last time the host looked at the device
    out->wait_sums_ms = (double)ctx->h_sc->wait_sums_ns * 1e-6;
    out->peer_memory = ctx->dist.p2p ? 1 : 0;
    out->persistent_loop = (ctx->has_state && !ctx->multi.on && loop_wanted(ctx) && ctx->loop_blocks_max > 0) ? 1 : 0;
    out->tile_lists = ctx->tile_valid ? 1 : 0;
    out->force_atoms_ms = (double)ctx->h_sc->force_atoms_ns * 1e-6;
    out->force_tail_ms = (double)ctx->h_sc->force_tail_ns * 1e-6;
    out->rebuild_ms = ctx->rebuild_host_ms;
    out->loop_steps = (int64_t)ctx->h_sc->loop_steps;
    for (int k = 0; k < 4; ++k) out->loop_phase_ms[k] = (double)ctx->h_sc->loop_ns[k] * 1e-6;
    return MD_OK;
}

void *md_stream(md_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int md_measure_fp64_peak(md_ctx *ctx, double *tflops)
{
    TRY(check_ctx(ctx, false));
    if (!tflops) return ctx->fail(MD_ERR_INVALID_ARGUMENT, "md_measure_fp64_peak: NULL output");
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
    double *sink = nullptr;
    TRY(dev_alloc(ctx, &sink, 1));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int iters = 1 << 15, blocks = sms * 8, threads = 256;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        CK(cudaEventRecord(e0, ctx->stream));
        k_fp64_peak<<<blocks, threads, 0, ctx->stream>>>(iters, 1.0 + rep, sink);
        CK(cudaEventRecord(e1, ctx->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double flop = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
        if (rep > 0 && ms > 0.f) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    dev_free(ctx, sink);
    *tflops = best;
    return MD_OK;
}

// MD_LOOP_TRACE builds: %globaltimer stamps (ns) of the last step the persistent loop ran — block 0: [0] step start,
// [1] drift done, [2] mid-step barrier passed, [3] forces done, [4] block sums, [5] ticket taken, [6] next step may start;
// last block: [8] enters the epilogue, [9] partials folded, [10] finalize done, [11] sequence number released.
__attribute__((visibility("default"))) int md_debug_trace(md_ctx *ctx, unsigned long long out[16])
{
    TRY(check_ctx(ctx, false));
    TRY(pull_scalars(ctx));
    for (int k = 0; k < 16; ++k) out[k] = ctx->h_sc->trace[k];
    return MD_OK;
}

int md_synchronize(md_ctx *ctx)
{
    TRY(check_ctx(ctx, false));
    CK(cudaStreamSynchronize(ctx->stream));
    return MD_OK;
}

int md_invalidate_lists(md_ctx *ctx)
{
    TRY(check_ctx(ctx, false));
    ctx->list_valid = false;
    return MD_OK;
}

}  // extern "C"

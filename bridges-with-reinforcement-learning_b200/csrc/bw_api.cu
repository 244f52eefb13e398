// C ABI of bridges_b200 (include/bridges_b200.h): handle management, state allocation in HBM,
// host<->device staging for the *_host entry points and kernel launches.  No torch types.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "bw_common.cuh"
#include "bw_kernels.cuh"
#include "bw_lp.cuh"

using namespace bw;

struct bw_handle {
    bw_config cfg;
    Params P;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    bool shapes_loaded = false;
    bool lp_enabled = false;   // stored bases allocated and BW_NO_LP not set
    bool timing = false;
    bool resets_unchecked = false;   // a bw_reset with tasks ran since the last look at P.reset_err
    bool force_staged = false;   // bw_set_host_transfer(h, 1): *_host calls always stage through device buffers
    cudaEvent_t ev[2] = {nullptr, nullptr};
    int smem_step = 0;
    int n_sm = 148, smem_per_sm = 233472;
    int64_t launches = 0;
    char err[512] = {0};
    // owned device buffers
    std::vector<void *> allocs;
    double *d_xs = nullptr, *d_ys = nullptr;
    ShapeDev *d_shapes = nullptr;
    // staging for *_host calls and internal evaluations
    bw_action *d_actions = nullptr, *d_noop = nullptr;
    bw_step_out *d_out = nullptr, *d_scratch_out = nullptr;
    uint8_t *d_mask = nullptr;
    bw_task *d_tasks = nullptr;
    float *d_img[3] = {nullptr, nullptr, nullptr};
    float *d_binary = nullptr;
    uint8_t *d_img_u8 = nullptr;
    uint64_t *d_bits_out = nullptr;      // staged copy of bw_obs_out.block_bits
    bw_interface *d_itf = nullptr;
    int32_t *d_nitf = nullptr;
    double *d_ground = nullptr, *d_offsets = nullptr;
    double ground_cached[256], offsets_cached[256];   // host copies of what d_ground / d_offsets hold
    int n_ground_cached = -1, n_offsets_cached = -1;
    // cache of candidate placements used by bw_enumerate_actions (bw_actions.cu, enumerate_kernel<true>)
    CandCache cand;
    int n_groups = 0;                // (shape, face) pairs with the target_faces bit: candidate groups
    bool cand_dirty = true;          // library or offset tables changed: every slot is stale
    size_t cand_budget = (size_t)4096 << 20;   // bytes; BW_CAND_CACHE_MB overrides, 0 disables the cache
    // fused rollout (bw_rollout_*): candidate buffers and the arguments of generate_actions
    RolloutBufs roll;
    bool roll_configured = false;
    double roll_ground[256], roll_offsets[256];
    int roll_n_ground = 0, roll_n_offsets = 0;
    int32_t roll_step = 0;
    bw_block *d_qblocks = nullptr, *d_rblocks = nullptr;
    uint8_t *d_qflags = nullptr;
    ShapeDev *d_rshapes = nullptr;
    uint64_t *d_rbits = nullptr;
    double *d_rxs = nullptr, *d_rys = nullptr;
};

// Shared-memory layout of the step kernel.  A launch that fits the GPU in one wave keeps everything in shared memory
// (two packed matrices, block library, pixel nodes): its length is the slowest environment's solve, nothing else
// matters.  With more environments than CTA slots the kernel is bound by how many latency-bound CTAs an SM holds:
// the two problems then share one packed matrix and the library stays in global memory when that buys a slot.
// BW_SHARE_H=0/1 overrides the choice (tuning hook, tools/ only).
static void choose_step_layout(bw_handle *h, int n_shapes) {
    Params &P = h->P;
    auto slots = [&](int bytes) {
        const int per_cta = bytes + 2560 /* static */ + 1024 /* reserved per CTA */;
        int c = h->smem_per_sm / per_cta;
        if (c > 8) c = 8;                        // 128 registers x 64 threads
        return c * h->n_sm;
    };
    const int full = step_smem_bytes(P.max_blocks, P.max_itf, n_shapes, false, true);
    bool share = false, lib = true;
    // (measured, profiles/README.md round 2: at 1.4 waves the turn-taking costs more than the extra slots give --
    // 6.42 against 6.73 M env steps/s on the bridge task at 1024 environments -- from two waves on it pays:
    // 2.19 against 2.69 ms per pass of the 65,536-assembly sweep)
    if (P.E >= 2 * slots(full)) {
        const int lean = step_smem_bytes(P.max_blocks, P.max_itf, n_shapes, true, true);
        const int leaner = step_smem_bytes(P.max_blocks, P.max_itf, n_shapes, true, false);
        if (slots(lean) > slots(full)) share = true;
        if (slots(leaner) > slots(lean)) { share = true; lib = false; }
    }
    if (const char *ov = getenv("BW_SHARE_H")) {
        share = atoi(ov) != 0;
        lib = !(share && atoi(ov) >= 2);
    }
    P.share_h = share ? 1 : 0;
    P.lib_in_smem = lib ? 1 : 0;
    h->smem_step = step_smem_bytes(P.max_blocks, P.max_itf, n_shapes, share, lib);
    // the LP path works in the memory of the two Newton problems (they never run at the same time)
    P.lp_on = (h->lp_enabled && lp_bytes(P.max_blocks, P.max_itf) <= step_problem_bytes(P.max_blocks, P.max_itf, share)) ? 1 : 0;
}

static int fail(bw_handle *h, int code, const char *fmt, ...) {
    if (h) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(h->err, sizeof(h->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) return fail(h, BW_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(_e)); \
    } while (0)

template <typename T>
static cudaError_t dev_alloc(bw_handle *h, T **p, size_t count, bool zero = true) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(T) > 0 ? count * sizeof(T) : 16);
    if (e != cudaSuccess) return e;
    h->allocs.push_back(q);
    if (zero) {
        e = cudaMemsetAsync(q, 0, count * sizeof(T), h->stream);
        if (e != cudaSuccess) return e;
    }
    *p = static_cast<T *>(q);
    return cudaSuccess;
}

// numpy.linspace(start, stop, num) with endpoint: arange(num) * step + start, last = stop
static void np_linspace(double start, double stop, int num, double *out) {
    const double step = (stop - start) / (double)(num - 1);
    for (int i = 0; i < num; i++) {
        volatile double prod = (double)i * step;   // separately rounded, as numpy does
        out[i] = prod + start;
    }
    out[num - 1] = stop;
}

static void shape_to_dev(const bw_shape_desc &s, ShapeDev &d) {
    memset(&d, 0, sizeof(d));
    d.n_faces = s.n_faces;
    d.n_verts = s.n_verts;
    d.target_faces_mask = s.target_faces_mask;
    d.receiving_faces_mask = s.receiving_faces_mask;
    for (int i = 0; i < BW_MAX_FACES; i++) {
        d.face_nx[i] = s.face_nx[i]; d.face_nz[i] = s.face_nz[i];
        d.face_cx[i] = s.face_cx[i]; d.face_cz[i] = s.face_cz[i];
        d.end0_x[i] = s.end0_x[i]; d.end0_z[i] = s.end0_z[i];
        d.end1_x[i] = s.end1_x[i]; d.end1_z[i] = s.end1_z[i];
    }
    double rad = 0.0;
    for (int i = 0; i < BW_MAX_VERTS; i++) {
        d.vert_x[i] = s.vert_x[i]; d.vert_z[i] = s.vert_z[i];
        if (i < s.n_verts) {
            const double dx = s.vert_x[i] - s.com_x, dz = s.vert_z[i] - s.com_z;
            rad = std::fmax(rad, std::sqrt(dx * dx + dz * dz));
        }
    }
    d.com_x = s.com_x; d.com_z = s.com_z; d.area = s.area; d.depth = s.depth;
    d.radius = rad;
}

static bool shape_ok(const bw_shape_desc &s) {
    return s.n_faces >= 3 && s.n_faces <= BW_MAX_FACES && s.n_verts >= 3 && s.n_verts <= BW_MAX_VERTS &&
           s.area > 0.0 && s.depth > 0.0;
}

// axis-aligned box marker (cube06.urdf: <box size="0.6 0.6 0.6">) in compas' face order
static void default_marker(ShapeDev &d) {
    bw_shape_desc s;
    memset(&s, 0, sizeof(s));
    const double hx = 0.5 * 0.6, hz = 0.5 * 0.6;
    s.n_faces = 4; s.n_verts = 4;
    s.target_faces_mask = s.receiving_faces_mask = 0xf;
    const double nx[4] = {0, 1, -1, 0}, nz[4] = {-1, 0, 0, 1};
    for (int i = 0; i < 4; i++) {
        s.face_nx[i] = nx[i]; s.face_nz[i] = nz[i];
        s.face_cx[i] = nx[i] * hx; s.face_cz[i] = nz[i] * hz;
    }
    s.end0_x[0] = -hx; s.end0_z[0] = -hz; s.end1_x[0] = hx; s.end1_z[0] = -hz;
    s.end0_x[1] = hx; s.end0_z[1] = -hz; s.end1_x[1] = hx; s.end1_z[1] = hz;
    s.end0_x[2] = -hx; s.end0_z[2] = hz; s.end1_x[2] = -hx; s.end1_z[2] = -hz;
    s.end0_x[3] = -hx; s.end0_z[3] = hz; s.end1_x[3] = hx; s.end1_z[3] = hz;
    const double vx[4] = {hx, -hx, -hx, hx}, vz[4] = {-hz, -hz, hz, hz};
    for (int i = 0; i < 4; i++) { s.vert_x[i] = vx[i]; s.vert_z[i] = vz[i]; }
    s.area = 0.36; s.depth = 0.6;
    shape_to_dev(s, d);
}

// Sizes (and, when the block library or the offset tables changed, clears) the candidate cache of
// bw_enumerate_actions.  Layout: per environment groups x (n_ground + max_blocks * NF * n_offsets) slots of
// 512 B raster + 4 B flags.  A layout that does not fit the budget -- or the device memory that is left --
// switches the cache off (plain kernel).
static void release_cand_cache(bw_handle *h) {
    CandCache &c = h->cand;
    void *old[7] = {c.meta, c.bits, c.pose, c.shape, c.seen_block, c.seen_obst, c.call};
    for (void *q : old) {
        if (!q) continue;
        cudaFree(q);
        for (size_t i = 0; i < h->allocs.size(); i++)
            if (h->allocs[i] == q) { h->allocs.erase(h->allocs.begin() + i); break; }
    }
    c = CandCache();
}

static int prepare_cand_cache(bw_handle *h, int n_ground, int n_offsets) {
    const int spg = n_ground + h->P.max_blocks * NF * n_offsets;
    const int slots = h->n_groups * spg;
    const size_t E = (size_t)h->P.E;
    const size_t need = E * (size_t)slots * (IMG * sizeof(uint64_t) + sizeof(uint32_t)) + E * NB * (sizeof(Pose) + 1) +
                        E * (2 * IMG * sizeof(uint64_t) + sizeof(uint32_t));
    CandCache &c = h->cand;
    if (slots <= 0 || need > h->cand_budget) {
        if (c.meta) {
            CU(cudaStreamSynchronize(h->stream));
            release_cand_cache(h);
        }
        return BW_OK;
    }
    if (c.meta == nullptr || c.slots != slots || c.spg != spg) {
        if (c.meta) {
            CU(cudaStreamSynchronize(h->stream));
            release_cand_cache(h);
        }
        cudaError_t e = dev_alloc(h, &c.meta, E * slots, false);
        if (e == cudaSuccess) e = dev_alloc(h, &c.bits, E * slots * IMG, false);
        if (e == cudaSuccess) e = dev_alloc(h, &c.pose, E * NB, false);
        if (e == cudaSuccess) e = dev_alloc(h, &c.shape, E * NB, false);
        if (e == cudaSuccess) e = dev_alloc(h, &c.seen_block, E * IMG, false);
        if (e == cudaSuccess) e = dev_alloc(h, &c.seen_obst, E * IMG, false);
        if (e == cudaSuccess) e = dev_alloc(h, &c.call, E, false);
        if (e != cudaSuccess) {
            // no room for it next to the caller's own allocations: enumerate without the cache from now on
            cudaGetLastError();
            release_cand_cache(h);
            h->cand_budget = 0;
            return BW_OK;
        }
        c.slots = slots;
        c.spg = spg;
        h->cand_dirty = true;
    }
    if (h->cand_dirty) {
        CU(cudaMemsetAsync(c.meta, 0, sizeof(uint32_t) * E * slots, h->stream));
        CU(cudaMemsetAsync(c.pose, 0xff, sizeof(Pose) * E * NB, h->stream));      // no block has this pose
        CU(cudaMemsetAsync(c.shape, 0xff, E * NB, h->stream));
        CU(cudaMemsetAsync(c.seen_block, 0, sizeof(uint64_t) * E * IMG, h->stream));
        CU(cudaMemsetAsync(c.seen_obst, 0, sizeof(uint64_t) * E * IMG, h->stream));
        CU(cudaMemsetAsync(c.call, 0, sizeof(uint32_t) * E, h->stream));
        h->cand_dirty = false;
    }
    return BW_OK;
}

extern "C" {

static int upload_offset_tables(bw_handle *h, const double *h_x_discr_ground, int n_ground, const double *h_offset_values,
                                int n_offsets);

int bw_abi_version(void) { return BW_ABI_VERSION; }

void bw_config_default(bw_config *cfg) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->num_envs = 1;
    cfg->device = 0;
    cfg->max_steps = 0;
    cfg->xlim[0] = -3.0; cfg->xlim[1] = 7.0;
    cfg->ylim[0] = 0.0; cfg->ylim[1] = 10.0;
    cfg->floor_halfwidth = 5.0;
    cfg->floor_depth = 10.0;
    cfg->mu = 0.8;
    cfg->density = 1.0;
    cfg->tmax = 1e-6;
    cfg->amin = 1e-3;
    cfg->stable_tol = 1e-6;
    cfg->stream = nullptr;
    cfg->collision_mode = 0;
    cfg->collision_tol = 0.005;
    cfg->bounds_lo[0] = -3.0; cfg->bounds_lo[1] = -3.0; cfg->bounds_lo[2] = -1.0;
    cfg->bounds_hi[0] = 7.0; cfg->bounds_hi[1] = 7.0; cfg->bounds_hi[2] = 9.0;
}

const char *bw_last_error(const bw_handle *h) { return h ? h->err : "null handle"; }

int64_t bw_kernel_launches(const bw_handle *h) { return h ? h->launches : 0; }

void bw_destroy(bw_handle *h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void *p : h->allocs) cudaFree(p);
    for (auto &e : h->ev)
        if (e) cudaEventDestroy(e);
    if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int bw_create(const bw_config *cfg, bw_handle **out) {
    if (!cfg || !out) return BW_ERR_INVALID;
    *out = nullptr;
    bw_handle *h = new (std::nothrow) bw_handle();
    if (!h) return BW_ERR_INVALID;
    h->cfg = *cfg;
    *out = h;   // returned even on failure so that bw_last_error works; caller destroys it
    if (cfg->num_envs <= 0) return fail(h, BW_ERR_INVALID, "num_envs must be positive");
    if (!(cfg->xlim[1] > cfg->xlim[0]) || !(cfg->ylim[1] > cfg->ylim[0]))
        return fail(h, BW_ERR_INVALID, "empty raster window");
    if (!(cfg->mu >= 0.0) || !(cfg->density > 0.0)) return fail(h, BW_ERR_INVALID, "mu/density out of range");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(h, BW_ERR_CUDA, "no CUDA device available (%s): bridges_b200 has no CPU path",
                    cudaGetErrorString(ce));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(h, BW_ERR_INVALID, "device %d out of range", cfg->device);
    CU(cudaSetDevice(cfg->device));
    if (cfg->use_caller_stream) {
        h->stream = static_cast<cudaStream_t>(cfg->stream);
    } else {
        CU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
        h->own_stream = true;
    }
    for (auto &e : h->ev) CU(cudaEventCreate(&e));

    const int E = cfg->num_envs;
    Params &P = h->P;
    memset(&P, 0, sizeof(P));
    P.E = E;
    P.max_steps = cfg->max_steps;
    // capacity used for shared-memory sizing: max_steps blocks when given, else the ABI maximum
    P.max_blocks = (cfg->max_steps > 0 && cfg->max_steps < BW_MAX_BLOCKS) ? cfg->max_steps : BW_MAX_BLOCKS;
    P.max_itf = 3 * P.max_blocks < BW_MAX_INTERFACES ? 3 * P.max_blocks : BW_MAX_INTERFACES;
    P.xlim0 = cfg->xlim[0]; P.xlim1 = cfg->xlim[1]; P.ylim0 = cfg->ylim[0]; P.ylim1 = cfg->ylim[1];
    P.floor_halfwidth = cfg->floor_halfwidth; P.floor_depth = cfg->floor_depth;
    P.density = cfg->density; P.tmax = cfg->tmax; P.amin = cfg->amin;
    P.stable_tol = cfg->stable_tol > 0 ? cfg->stable_tol : 1e-6;
    P.inv_step_x = (double)(IMG - 1) / (cfg->xlim[1] - cfg->xlim[0]);
    P.inv_step_y = (double)(IMG - 1) / (cfg->ylim[1] - cfg->ylim[0]);

    double xs[IMG], ys[IMG];
    np_linspace(cfg->xlim[0], cfg->xlim[1], IMG, xs);
    np_linspace(cfg->ylim[1], cfg->ylim[0], IMG, ys);
    CU(dev_alloc(h, &h->d_xs, IMG));
    CU(dev_alloc(h, &h->d_ys, IMG));
    CU(cudaMemcpyAsync(h->d_xs, xs, sizeof(xs), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_ys, ys, sizeof(ys), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));   // xs/ys live on this stack frame
    P.xs = h->d_xs; P.ys = h->d_ys;
    CU(dev_alloc(h, &h->d_shapes, BW_MAX_SHAPES + 1));    // the last entry is the obstacle / target marker
    P.shapes = h->d_shapes;
    P.marker = h->d_shapes + BW_MAX_SHAPES;
    if (cfg->collision_mode != 0 && cfg->collision_mode != 1) return fail(h, BW_ERR_INVALID, "collision_mode must be 0 or 1");
    P.collision_mode = cfg->collision_mode;
    P.screen = getenv("BW_NO_SCREEN") ? 0 : 1;     // tuning hook (tools/ only): solver without the mechanism screen
    if (const char *mb = getenv("BW_CAND_CACHE_MB")) h->cand_budget = (size_t)strtoull(mb, nullptr, 10) << 20;
    P.collision_tol = cfg->collision_tol;
    for (int k = 0; k < 3; k++) { P.bounds_lo[k] = cfg->bounds_lo[k]; P.bounds_hi[k] = cfg->bounds_hi[k]; }

    CU(dev_alloc(h, &P.n_blocks, E));
    CU(dev_alloc(h, &P.pose, (size_t)E * NB));
    CU(dev_alloc(h, &P.shape_of, (size_t)E * NB));
    CU(dev_alloc(h, &P.face_occ, (size_t)E * NB));
    CU(dev_alloc(h, &P.static_mask, E));
    CU(dev_alloc(h, &P.block_bits, (size_t)E * IMG));
    CU(dev_alloc(h, &P.obst_bits, (size_t)E * IMG));
    CU(dev_alloc(h, &P.reward_img, (size_t)E * IMG * IMG));
    CU(dev_alloc(h, &P.task, E));
    CU(dev_alloc(h, &P.mu, E, false));
    CU(dev_alloc(h, &P.done, E));
    CU(dev_alloc(h, &P.last_out, E));
    CU(dev_alloc(h, &P.su_valid, E));
    CU(dev_alloc(h, &P.warm_y, (size_t)E * 2 * NB * 3));
    CU(dev_alloc(h, &P.warm_ok, (size_t)E * 2));
    P.warm_start = getenv("BW_NO_WARM") ? 0 : 1;   // tuning hook (tools/ only): every solve from y = 0
    {   // stored bases of the LP verdict path (bw_lp.cuh); BW_NO_LP: tuning hook (tools/ only)
        const int mm = 3 * P.max_blocks;
        P.lp_stride = mm * lp_row_stride(mm);
        CU(dev_alloc(h, &P.lp_meta, E));
        CU(dev_alloc(h, &P.lp_binv, (size_t)E * P.lp_stride, false));
        CU(dev_alloc(h, &P.lp_ids, (size_t)E * 3 * NB, false));
        CU(dev_alloc(h, &P.lp_xb, (size_t)E * 3 * NB, false));
        h->lp_enabled = getenv("BW_NO_LP") == nullptr;
        P.lp_par = getenv("BW_LP_SEQ") ? 0 : 1;        // tuning hook (tools/ only): the two problems one after the other
        if (getenv("BW_LP_STATS")) CU(dev_alloc(h, &P.lp_stats, 32));   // tuning hook (tools/ only)
    }
    CU(dev_alloc(h, &P.cand_need, 1));
    CU(dev_alloc(h, &P.reset_err, 1));
    {   // CTA order of the step kernel: the first launch takes the environments in index order
        CU(dev_alloc(h, &P.order_cnt, 3 * ORDER_KEYS));
        CU(dev_alloc(h, &P.order_q, (size_t)3 * ORDER_KEYS * E, false));
        std::vector<int32_t> ident(E);
        for (int i = 0; i < E; i++) ident[i] = i;
        const int32_t all = E;
        CU(cudaMemcpyAsync(P.order_q, ident.data(), sizeof(int32_t) * E, cudaMemcpyHostToDevice, h->stream));   // queue 0, class 0
        CU(cudaMemcpyAsync(P.order_cnt, &all, sizeof(int32_t), cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
        P.order_phase = 0;
        P.order_on = getenv("BW_NO_ORDER") ? 0 : 1;   // tuning hook (tools/ only): CTA i = environment i
    }
    {
        std::vector<double> mu(E, cfg->mu);
        CU(cudaMemcpyAsync(P.mu, mu.data(), sizeof(double) * E, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    CU(dev_alloc(h, &h->d_actions, E));
    CU(dev_alloc(h, &h->d_noop, E, false));
    {
        std::vector<bw_action> noop(E);
        memset(noop.data(), 0, sizeof(bw_action) * E);
        for (auto &a : noop) { a.shape = -1; a.target_block = -1; }
        CU(cudaMemcpyAsync(h->d_noop, noop.data(), sizeof(bw_action) * E, cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    CU(dev_alloc(h, &h->d_out, E));
    CU(dev_alloc(h, &h->d_scratch_out, E));
    CU(dev_alloc(h, &h->d_mask, E));
    CU(dev_alloc(h, &h->d_ground, 256));
    CU(dev_alloc(h, &h->d_offsets, 256));

    upload_step_tables();
    // Gaussian of get_task_features (kernel_size 101, sigma 16), float32 like torch
    {
        float k[127];
        double sum = 0.0;
        for (int i = 0; i < 101; i++) {
            const float c = (float)(i - 50);
            k[i] = expf(-(c * c) / 512.0f);
            sum += k[i];
        }
        const float fs = (float)sum;
        for (int i = 0; i < 101; i++) k[i] = k[i] / fs;
        ShapeDev marker;
        default_marker(marker);
        upload_obs_tables(k, &marker);
        CU(cudaMemcpyAsync(h->d_shapes + BW_MAX_SHAPES, &marker, sizeof(ShapeDev), cudaMemcpyHostToDevice, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    // the step kernel keeps the block library in shared memory: sized for the largest library here,
    // re-sized for the actual one by bw_load_shapes
    {
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, cfg->device));
        h->n_sm = prop.multiProcessorCount;
        h->smem_per_sm = (int)prop.sharedMemPerMultiprocessor;
    }
    choose_step_layout(h, BW_MAX_SHAPES);
    CU(configure_step(step_smem_bytes(P.max_blocks, P.max_itf, BW_MAX_SHAPES, false, true)));   // the largest layout
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(h->stream));
    return BW_OK;
}

// tasks refused by reset_kernel since the last check (the kernel cannot return a status itself)
static int check_reset_errors(bw_handle *h) {
    int32_t bad = 0;
    CU(cudaMemcpyAsync(&bad, h->P.reset_err, sizeof(bad), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    if (bad == 0) return BW_OK;
    CU(cudaMemsetAsync(h->P.reset_err, 0, sizeof(int32_t), h->stream));
    return fail(h, BW_ERR_INVALID, "reset: %d task(s) refused (pre-placed blocks exceed max_steps / BW_MAX_BLOCKS, "
                                   "or a shape index / obstacle / target count out of range); those environments were left empty",
                (int)bad);
}

int bw_sync(bw_handle *h) {
    if (!h) return BW_ERR_INVALID;
    CU(cudaStreamSynchronize(h->stream));
    if (h->resets_unchecked) {
        h->resets_unchecked = false;
        return check_reset_errors(h);
    }
    return BW_OK;
}

int bw_load_shapes(bw_handle *h, const bw_shape_desc *h_shapes, int32_t n) {
    if (!h || !h_shapes) return BW_ERR_INVALID;
    if (n <= 0 || n > BW_MAX_SHAPES) return fail(h, BW_ERR_CAPACITY, "1..%d shapes supported, got %d", BW_MAX_SHAPES, n);
    ShapeDev dev[BW_MAX_SHAPES];
    for (int i = 0; i < n; i++) {
        if (!shape_ok(h_shapes[i])) return fail(h, BW_ERR_INVALID, "shape %d: bad face/vertex count or mass data", i);
        shape_to_dev(h_shapes[i], dev[i]);
    }
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(h->d_shapes, dev, sizeof(ShapeDev) * n, cudaMemcpyHostToDevice, h->stream));
    // verdicts of the last step describe blocks of the old library
    CU(cudaMemsetAsync(h->P.su_valid, 0, h->P.E, h->stream));
    CU(cudaMemsetAsync(h->P.warm_ok, 0, (size_t)h->P.E * 2, h->stream));
    CU(cudaMemsetAsync(h->P.lp_meta, 0, (size_t)h->P.E * sizeof(LpMeta), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    h->P.n_shapes = n;
    choose_step_layout(h, n);
    h->shapes_loaded = true;
    h->n_groups = 0;
    for (int i = 0; i < n; i++)
        for (int f = 0; f < dev[i].n_faces; f++)
            if ((dev[i].target_faces_mask >> f) & 1u) h->n_groups++;
    h->cand_dirty = true;
    return BW_OK;
}

int bw_set_marker_shape(bw_handle *h, const bw_shape_desc *h_shape) {
    if (!h || !h_shape) return BW_ERR_INVALID;
    if (!shape_ok(*h_shape)) return fail(h, BW_ERR_INVALID, "marker shape: bad face/vertex count or mass data");
    ShapeDev d;
    shape_to_dev(*h_shape, d);
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    upload_obs_tables(nullptr, &d);
    CU(cudaMemcpyAsync(h->d_shapes + BW_MAX_SHAPES, &d, sizeof(ShapeDev), cudaMemcpyHostToDevice, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return BW_OK;
}

int bw_set_task_kernel(bw_handle *h, const float *h_kernel1d, int32_t n) {
    if (!h || !h_kernel1d) return BW_ERR_INVALID;
    if (n != 101) return fail(h, BW_ERR_INVALID, "the task-feature kernel has 101 taps (successor_dqn.py:78)");
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaStreamSynchronize(h->stream));
    upload_obs_tables(h_kernel1d, nullptr);
    CU(cudaGetLastError());
    return BW_OK;
}

int bw_set_mu(bw_handle *h, const double *h_mu) {
    if (!h || !h_mu) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(h->P.mu, h_mu, sizeof(double) * h->P.E, cudaMemcpyHostToDevice, h->stream));
    // the released-block verdict of the last step was computed with the old coefficients: it must not
    // stand in for the next step's frozen solve (step_kernel, prev_released_ok)
    CU(cudaMemsetAsync(h->P.su_valid, 0, h->P.E, h->stream));
    CU(cudaMemsetAsync(h->P.warm_ok, 0, (size_t)h->P.E * 2, h->stream));
    CU(cudaMemsetAsync(h->P.lp_meta, 0, (size_t)h->P.E * sizeof(LpMeta), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BW_OK;
}

int bw_set_static_mask(bw_handle *h, const uint32_t *h_mask) {
    if (!h || !h_mask) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(h->P.static_mask, h_mask, sizeof(uint32_t) * h->P.E, cudaMemcpyHostToDevice, h->stream));
    // the verdicts of the last step no longer describe these supports
    CU(cudaMemsetAsync(h->P.su_valid, 0, h->P.E, h->stream));
    CU(cudaMemsetAsync(h->P.warm_ok, 0, (size_t)h->P.E * 2, h->stream));
    CU(cudaMemsetAsync(h->P.lp_meta, 0, (size_t)h->P.E * sizeof(LpMeta), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BW_OK;
}

int bw_debug_lp_stats(bw_handle *h, uint64_t *h_stats /*[32]*/) {
    if (!h || !h_stats) return BW_ERR_INVALID;
    memset(h_stats, 0, 32 * sizeof(uint64_t));
    if (h->P.lp_stats == nullptr) return BW_OK;
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(h_stats, h->P.lp_stats, 32 * sizeof(uint64_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BW_OK;
}

int bw_fp64_peak_gflops(bw_handle *h, double *h_gflops) {
    if (!h || !h_gflops) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    *h_gflops = measure_fp64_gflops(h->stream);
    h->launches += 4;
    CU(cudaGetLastError());
    return BW_OK;
}

int bw_set_timing(bw_handle *h, int32_t enabled) {
    if (!h) return BW_ERR_INVALID;
    h->timing = enabled != 0;
    return BW_OK;
}

static int need_shapes(bw_handle *h) {
    if (!h->shapes_loaded) return fail(h, BW_ERR_STATE, "bw_load_shapes must be called first");
    return BW_OK;
}

int bw_reset(bw_handle *h, const bw_task *d_tasks, const uint8_t *d_mask) {
    if (!h) return BW_ERR_INVALID;
    if (int rc = need_shapes(h)) return rc;
    CU(cudaSetDevice(h->cfg.device));
    launch_reset(h->P, d_tasks, d_mask, 0, h->stream);
    h->launches++;
    if (d_tasks != nullptr) {
        h->resets_unchecked = true;
        // pre-placed blocks: refresh the verdicts / distances as add_block does (gym_env.py:279-281)
        launch_step(h->P, nullptr, d_mask, h->d_scratch_out, bw_obs_out{nullptr, nullptr, nullptr, nullptr}, nullptr, nullptr,
                    0, h->smem_step, h->stream);
        h->launches++;
    }
    CU(cudaGetLastError());
    return BW_OK;
}

int bw_reset_host(bw_handle *h, const bw_task *h_tasks, const uint8_t *h_mask) {
    if (!h) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    const int E = h->P.E;
    if (h_tasks) {
        for (int e = 0; e < E; e++) {
            if (h_mask && !h_mask[e]) continue;
            const bw_task &t = h_tasks[e];
            if (t.n_blocks < 0 || t.n_blocks > h->P.max_blocks)
                return fail(h, BW_ERR_CAPACITY, "task %d: %d pre-placed blocks, this handle holds %d (max_steps / BW_MAX_BLOCKS)",
                            e, t.n_blocks, h->P.max_blocks);
            if (t.n_obstacles < 0 || t.n_obstacles > BW_MAX_OBSTACLES || t.n_targets < 0 || t.n_targets > BW_MAX_TARGETS)
                return fail(h, BW_ERR_CAPACITY, "task %d: obstacle / target count out of range", e);
            for (int i = 0; i < t.n_blocks; i++)
                if (t.blocks[i].shape < 0 || t.blocks[i].shape >= h->P.n_shapes)
                    return fail(h, BW_ERR_INVALID, "task %d: pre-placed block %d has shape index %d, the library has %d shapes",
                                e, i, t.blocks[i].shape, h->P.n_shapes);
        }
        if (!h->d_tasks) CU(dev_alloc(h, &h->d_tasks, E));
        CU(cudaMemcpyAsync(h->d_tasks, h_tasks, sizeof(bw_task) * E, cudaMemcpyHostToDevice, h->stream));
    }
    if (h_mask) CU(cudaMemcpyAsync(h->d_mask, h_mask, E, cudaMemcpyHostToDevice, h->stream));
    int rc = bw_reset(h, h_tasks ? h->d_tasks : nullptr, h_mask ? h->d_mask : nullptr);
    if (rc) return rc;
    return bw_sync(h);
}

int bw_reset_done(bw_handle *h) {
    if (!h) return BW_ERR_INVALID;
    if (int rc = need_shapes(h)) return rc;
    CU(cudaSetDevice(h->cfg.device));
    launch_reset(h->P, nullptr, nullptr, 1, h->stream);
    h->launches++;
    CU(cudaGetLastError());
    return BW_OK;
}

int bw_step(bw_handle *h, const bw_action *d_actions, const uint8_t *d_mask, bw_step_out *d_out,
            const bw_obs_out *obs) {
    if (!h || !d_actions || !d_out) return BW_ERR_INVALID;
    if (int rc = need_shapes(h)) return rc;
    CU(cudaSetDevice(h->cfg.device));
    if (h->timing) CU(cudaEventRecord(h->ev[0], h->stream));
    // one kernel: placement, interfaces, both solves, bookkeeping, raster update and the
    // observation write
    launch_step(h->P, d_actions, d_mask, d_out, obs ? *obs : bw_obs_out{nullptr, nullptr, nullptr, nullptr}, nullptr, nullptr, 0,
                h->smem_step, h->stream);
    h->launches++;
    if (h->timing) CU(cudaEventRecord(h->ev[1], h->stream));
    CU(cudaGetLastError());
    return BW_OK;
}

int bw_evaluate(bw_handle *h, const uint8_t *d_mask, bw_step_out *d_out, const bw_obs_out *obs) {
    if (!h || !d_out) return BW_ERR_INVALID;
    if (int rc = need_shapes(h)) return rc;
    CU(cudaSetDevice(h->cfg.device));
    if (h->timing) CU(cudaEventRecord(h->ev[0], h->stream));
    // the step kernel without an action: interfaces, both verdicts, distances, observations of what stands
    launch_step(h->P, nullptr, d_mask, d_out, obs ? *obs : bw_obs_out{nullptr, nullptr, nullptr, nullptr}, nullptr, nullptr, 0,
                h->smem_step, h->stream);
    h->launches++;
    if (h->timing) CU(cudaEventRecord(h->ev[1], h->stream));
    CU(cudaGetLastError());
    return BW_OK;
}

int bw_last_step_kernel_ms(bw_handle *h, float *h_ms2) {
    if (!h || !h_ms2) return BW_ERR_INVALID;
    if (!h->timing) return fail(h, BW_ERR_STATE, "bw_set_timing(h, 1) first");
    CU(cudaStreamSynchronize(h->stream));
    h_ms2[0] = h_ms2[1] = 0.0f;
    CU(cudaEventElapsedTime(&h_ms2[0], h->ev[0], h->ev[1]));
    return BW_OK;
}

static int ensure_img(bw_handle *h, int which) {
    if (!h->d_img[which]) CU(dev_alloc(h, &h->d_img[which], (size_t)h->P.E * IMG * IMG, false));
    return BW_OK;
}

// Device-visible alias of a pinned (page-locked, mapped) host buffer, or nullptr for pageable memory.
// With unified addressing every cudaHostAlloc / cudaHostRegister'ed range has one (torch's
// pin_memory() included).
static void *mapped_alias(const void *host_ptr) {
    if (!host_ptr) return nullptr;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, host_ptr) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    if (attr.type != cudaMemoryTypeHost || attr.devicePointer == nullptr) return nullptr;
    return attr.devicePointer;
}

int bw_step_host(bw_handle *h, const bw_action *h_actions, const uint8_t *h_mask, bw_step_out *h_out,
                 const bw_obs_out *obs) {
    if (!h || !h_actions || !h_out) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    const size_t E = (size_t)h->P.E;
    // Zero-copy path: when every host buffer of this call is pinned, the step kernel reads the
    // actions from and writes its records / images straight to host memory over PCIe, each
    // environment as soon as it is finished -- the transfers overlap the solves of the slower
    // environments instead of following the kernel.  Pageable buffers take the staged path below.
    {
        void *m_act = mapped_alias(h_actions), *m_out = mapped_alias(h_out);
        void *m_mask = h_mask ? mapped_alias(h_mask) : nullptr;
        void *m_f32 = (obs && obs->block_img_f32) ? mapped_alias(obs->block_img_f32) : nullptr;
        void *m_u8 = (obs && obs->block_img_u8) ? mapped_alias(obs->block_img_u8) : nullptr;
        void *m_bin = (obs && obs->binary) ? mapped_alias(obs->binary) : nullptr;
        void *m_bits = (obs && obs->block_bits) ? mapped_alias(obs->block_bits) : nullptr;
        const bool all_mapped = m_act && m_out && (!h_mask || m_mask) && (!(obs && obs->block_img_f32) || m_f32) &&
                                (!(obs && obs->block_img_u8) || m_u8) && (!(obs && obs->binary) || m_bin) &&
                                (!(obs && obs->block_bits) || m_bits);
        if (all_mapped && !h->force_staged) {
            bw_obs_out dev = {static_cast<float *>(m_f32), static_cast<uint8_t *>(m_u8), static_cast<float *>(m_bin),
                              static_cast<uint64_t *>(m_bits)};
            int rc = bw_step(h, static_cast<const bw_action *>(m_act), static_cast<const uint8_t *>(m_mask),
                             static_cast<bw_step_out *>(m_out), &dev);
            if (rc) return rc;
            CU(cudaStreamSynchronize(h->stream));
            return BW_OK;
        }
    }
    CU(cudaMemcpyAsync(h->d_actions, h_actions, sizeof(bw_action) * E, cudaMemcpyHostToDevice, h->stream));
    if (h_mask) CU(cudaMemcpyAsync(h->d_mask, h_mask, E, cudaMemcpyHostToDevice, h->stream));
    bw_obs_out dev = {nullptr, nullptr, nullptr, nullptr};
    if (obs && obs->block_img_f32) {
        if (int rc = ensure_img(h, 0)) return rc;
        dev.block_img_f32 = h->d_img[0];
    }
    if (obs && obs->block_img_u8) {
        if (!h->d_img_u8) CU(dev_alloc(h, &h->d_img_u8, E * IMG * IMG, false));
        dev.block_img_u8 = h->d_img_u8;
    }
    if (obs && obs->binary) {
        if (!h->d_binary) CU(dev_alloc(h, &h->d_binary, E * 6, false));
        dev.binary = h->d_binary;
    }
    if (obs && obs->block_bits) {
        if (!h->d_bits_out) CU(dev_alloc(h, &h->d_bits_out, E * IMG, false));
        dev.block_bits = h->d_bits_out;
    }
    int rc = bw_step(h, h->d_actions, h_mask ? h->d_mask : nullptr, h->d_out, &dev);
    if (rc) return rc;
    CU(cudaMemcpyAsync(h_out, h->d_out, sizeof(bw_step_out) * E, cudaMemcpyDeviceToHost, h->stream));
    if (dev.block_img_f32)
        CU(cudaMemcpyAsync(obs->block_img_f32, dev.block_img_f32, sizeof(float) * E * IMG * IMG, cudaMemcpyDeviceToHost,
                           h->stream));
    if (dev.block_img_u8)
        CU(cudaMemcpyAsync(obs->block_img_u8, dev.block_img_u8, E * IMG * IMG, cudaMemcpyDeviceToHost, h->stream));
    if (dev.binary)
        CU(cudaMemcpyAsync(obs->binary, dev.binary, sizeof(float) * E * 6, cudaMemcpyDeviceToHost, h->stream));
    if (dev.block_bits)
        CU(cudaMemcpyAsync(obs->block_bits, dev.block_bits, sizeof(uint64_t) * E * IMG, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BW_OK;
}

int bw_set_host_transfer(bw_handle *h, int32_t mode) {
    if (!h || mode < 0 || mode > 1) return BW_ERR_INVALID;
    h->force_staged = (mode == 1);
    return BW_OK;
}

int bw_observe(bw_handle *h, float *d_block_img, float *d_binary, float *d_obstacle_img, float *d_reward_img) {
    if (!h) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    launch_observe(h->P, d_block_img, d_binary, d_obstacle_img, d_reward_img, h->stream);
    h->launches += (d_block_img ? 1 : 0) + (d_binary ? 1 : 0) + (d_obstacle_img ? 1 : 0) + (d_reward_img ? 1 : 0);
    CU(cudaGetLastError());
    return BW_OK;
}

int bw_observe_host(bw_handle *h, float *h_block_img, float *h_binary, float *h_obstacle_img, float *h_reward_img) {
    if (!h) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    const size_t E = (size_t)h->P.E, img_bytes = sizeof(float) * E * IMG * IMG;
    float *hosts[3] = {h_block_img, h_obstacle_img, h_reward_img};
    for (int i = 0; i < 3; i++)
        if (hosts[i])
            if (int rc = ensure_img(h, i)) return rc;
    if (h_binary && !h->d_binary) CU(dev_alloc(h, &h->d_binary, E * 6, false));
    int rc = bw_observe(h, h_block_img ? h->d_img[0] : nullptr, h_binary ? h->d_binary : nullptr,
                        h_obstacle_img ? h->d_img[1] : nullptr, h_reward_img ? h->d_img[2] : nullptr);
    if (rc) return rc;
    for (int i = 0; i < 3; i++)
        if (hosts[i]) CU(cudaMemcpyAsync(hosts[i], h->d_img[i], img_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (h_binary)
        CU(cudaMemcpyAsync(h_binary, h->d_binary, sizeof(float) * E * 6, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BW_OK;
}

static int enumerate_common(bw_handle *h, const double *h_x_discr_ground, int32_t n_ground, const double *h_offset_values,
                            int32_t n_offsets, int32_t amax, bw_action *d_cand, uint8_t *d_valid, int32_t *d_n_cand,
                            uint64_t *d_action_bits, int32_t *d_slot) {
    if (!h || !d_cand || !d_valid || !d_n_cand || amax <= 0) return BW_ERR_INVALID;
    if (int rc = need_shapes(h)) return rc;
    if (n_ground < 0 || n_ground > 256 || n_offsets < 0 || n_offsets > 256)
        return fail(h, BW_ERR_CAPACITY, "at most 256 ground offsets / block offsets");
    if ((n_ground > 0 && !h_x_discr_ground) || (n_offsets > 0 && !h_offset_values)) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    if (int rc = upload_offset_tables(h, h_x_discr_ground, n_ground, h_offset_values, n_offsets)) return rc;
    if (int rc = prepare_cand_cache(h, n_ground, n_offsets)) return rc;
    if (d_slot != nullptr && h->cand.meta == nullptr)
        return fail(h, BW_ERR_CAPACITY, "no candidate store (BW_CAND_CACHE_MB, device memory): use bw_enumerate_actions");
    launch_enumerate(h->P, h->d_ground, n_ground, h->d_offsets, n_offsets, amax, d_cand, d_valid, d_n_cand,
                     d_action_bits, d_slot, h->cand, h->stream);
    h->launches++;
    CU(cudaGetLastError());
    return BW_OK;
}

int bw_enumerate_actions(bw_handle *h, const double *h_x_discr_ground, int32_t n_ground, const double *h_offset_values,
                         int32_t n_offsets, int32_t amax, bw_action *d_cand, uint8_t *d_valid, int32_t *d_n_cand,
                         uint64_t *d_action_bits) {
    return enumerate_common(h, h_x_discr_ground, n_ground, h_offset_values, n_offsets, amax, d_cand, d_valid, d_n_cand,
                            d_action_bits, nullptr);
}

int bw_enumerate_actions_stored(bw_handle *h, const double *h_x_discr_ground, int32_t n_ground,
                                const double *h_offset_values, int32_t n_offsets, int32_t amax, bw_action *d_cand,
                                uint8_t *d_valid, int32_t *d_n_cand, int32_t *d_slot) {
    if (!d_slot) return BW_ERR_INVALID;
    return enumerate_common(h, h_x_discr_ground, n_ground, h_offset_values, n_offsets, amax, d_cand, d_valid, d_n_cand,
                            nullptr, d_slot);
}

int bw_gather_action_bits(bw_handle *h, const int32_t *d_slot, int32_t amax, const int32_t *d_env, const int32_t *d_index,
                          int64_t n, uint64_t *d_bits) {
    if (!h || !d_slot || !d_index || !d_bits || amax <= 0 || n < 0) return BW_ERR_INVALID;
    if (h->cand.meta == nullptr) return fail(h, BW_ERR_STATE, "no candidate store: nothing was enumerated into it");
    CU(cudaSetDevice(h->cfg.device));
    if (n > 0) {
        launch_gather_bits(h->cand, d_slot, nullptr, amax, h->P.E, d_env, d_index, n, d_bits, h->stream);
        h->launches++;
    }
    CU(cudaGetLastError());
    return BW_OK;
}

// the offset tables rarely change between calls: upload only when they do
static int upload_offset_tables(bw_handle *h, const double *h_x_discr_ground, int n_ground, const double *h_offset_values,
                                int n_offsets) {
    if (n_ground > 0 && (h->n_ground_cached != n_ground ||
                         memcmp(h->ground_cached, h_x_discr_ground, sizeof(double) * n_ground) != 0)) {
        memcpy(h->ground_cached, h_x_discr_ground, sizeof(double) * n_ground);
        h->n_ground_cached = n_ground;
        h->cand_dirty = true;
        CU(cudaMemcpyAsync(h->d_ground, h->ground_cached, sizeof(double) * n_ground, cudaMemcpyHostToDevice, h->stream));
    }
    if (n_offsets > 0 && (h->n_offsets_cached != n_offsets ||
                          memcmp(h->offsets_cached, h_offset_values, sizeof(double) * n_offsets) != 0)) {
        memcpy(h->offsets_cached, h_offset_values, sizeof(double) * n_offsets);
        h->n_offsets_cached = n_offsets;
        h->cand_dirty = true;
        CU(cudaMemcpyAsync(h->d_offsets, h->offsets_cached, sizeof(double) * n_offsets, cudaMemcpyHostToDevice, h->stream));
    }
    return BW_OK;
}

// ---- fused lock-step rollout -------------------------------------------------------------------------------
int bw_rollout_configure(bw_handle *h, const double *h_x_discr_ground, int32_t n_ground, const double *h_offset_values,
                         int32_t n_offsets, int32_t amax, int32_t env_id_base) {
    if (!h || amax <= 0) return BW_ERR_INVALID;
    if (int rc = need_shapes(h)) return rc;
    if (n_ground < 0 || n_ground > 256 || n_offsets < 0 || n_offsets > 256)
        return fail(h, BW_ERR_CAPACITY, "at most 256 ground offsets / block offsets");
    if ((n_ground > 0 && !h_x_discr_ground) || (n_offsets > 0 && !h_offset_values)) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    const size_t E = (size_t)h->P.E;
    RolloutBufs &R = h->roll;
    if (R.cand == nullptr || R.amax != amax) {
        if (R.cand != nullptr) {
            CU(cudaStreamSynchronize(h->stream));
            void *old[4] = {R.cand, R.valid, R.bits, R.slot};
            for (void *q : old) {
                if (!q) continue;
                cudaFree(q);
                for (size_t i = 0; i < h->allocs.size(); i++)
                    if (h->allocs[i] == q) { h->allocs.erase(h->allocs.begin() + i); break; }
            }
        }
        CU(dev_alloc(h, &R.cand, E * amax));
        CU(dev_alloc(h, &R.valid, E * amax));
        R.bits = nullptr;                 // dense raster copies: allocated by rollout_enumerate without a store
        CU(dev_alloc(h, &R.slot, E * amax));
        if (R.n_cand == nullptr) {
            CU(dev_alloc(h, &R.n_cand, E));
            CU(dev_alloc(h, &R.n_valid, E));
            CU(dev_alloc(h, &R.actions, E));
            CU(dev_alloc(h, &R.has_action, E));
            CU(dev_alloc(h, &R.stuck, E));
        }
        R.amax = amax;
    }
    R.env_id_base = env_id_base;
    if (n_ground > 0) memcpy(h->roll_ground, h_x_discr_ground, sizeof(double) * n_ground);
    if (n_offsets > 0) memcpy(h->roll_offsets, h_offset_values, sizeof(double) * n_offsets);
    h->roll_n_ground = n_ground;
    h->roll_n_offsets = n_offsets;
    h->roll_configured = true;
    return BW_OK;
}

// Candidates of the current states into the rollout buffers; environments left without any candidate (other than
// fresh ones) are restarted and enumerated once more.  With a candidate store this is ONE launch that closes the
// iteration (RollFuse): d_out given -> the record of the step that has just run + the restart of finished episodes
// come first; d_next_slots given -> the built-in random policy picks for the next iteration right away (*picked).
static int rollout_enumerate(bw_handle *h, bw_transition *d_slots, const bw_step_out *d_out, bw_transition *d_next_slots,
                             uint64_t seed, bool *picked) {
    RolloutBufs &R = h->roll;
    if (picked) *picked = false;
    if (int rc = upload_offset_tables(h, h->roll_ground, h->roll_n_ground, h->roll_offsets, h->roll_n_offsets)) return rc;
    if (int rc = prepare_cand_cache(h, h->roll_n_ground, h->roll_n_offsets)) return rc;
    // with a candidate store the rasters stay where they are (R.slot says where); without one they are copied out
    const bool stored = h->cand.meta != nullptr;
    if (stored) {
        RollFuse F;
        F.R = R;
        F.R.bits = nullptr;               // rasters are read out of the store
        F.slots = d_slots;
        F.out = d_out;
        F.next_slots = d_next_slots;
        F.seed = seed;
        F.next_step = h->roll_step;
        launch_enumerate(h->P, h->d_ground, h->roll_n_ground, h->d_offsets, h->roll_n_offsets, R.amax, R.cand, R.valid,
                         R.n_cand, nullptr, R.slot, h->cand, h->stream, nullptr, R.n_valid, &F);
        h->launches += 1;
        if (picked) *picked = d_next_slots != nullptr;
    } else {
        if (R.bits == nullptr) CU(dev_alloc(h, &R.bits, (size_t)h->P.E * R.amax * IMG, false));
        if (d_out != nullptr) {
            launch_rollout_record(h->P, R, d_out, d_slots, h->stream);     // finished episodes start afresh (their task is kept)
            h->launches += 1;
        }
        launch_enumerate(h->P, h->d_ground, h->roll_n_ground, h->d_offsets, h->roll_n_offsets, R.amax, R.cand, R.valid,
                         R.n_cand, R.bits, R.slot, h->cand, h->stream, nullptr, R.n_valid);
        launch_rollout_finalize(h->P, R, d_slots, h->stream);      // restarts the environments left without a candidate
        launch_enumerate(h->P, h->d_ground, h->roll_n_ground, h->d_offsets, h->roll_n_offsets, R.amax, R.cand, R.valid,
                         R.n_cand, R.bits, R.slot, h->cand, h->stream, R.stuck, R.n_valid);
        h->launches += 3;
    }
    CU(cudaGetLastError());
    return BW_OK;
}

int bw_rollout_begin(bw_handle *h, bw_rollout_view *out) {
    if (!h) return BW_ERR_INVALID;
    if (!h->roll_configured) return fail(h, BW_ERR_STATE, "bw_rollout_configure must be called first");
    CU(cudaSetDevice(h->cfg.device));
    if (int rc = rollout_enumerate(h, nullptr, nullptr, nullptr, 0, nullptr)) return rc;
    if (out) {
        const RolloutBufs &R = h->roll;
        out->cand = R.cand; out->valid = R.valid; out->n_cand = R.n_cand; out->n_valid = R.n_valid;
        out->action_bits = (h->cand.meta != nullptr) ? nullptr : R.bits;
        out->slot = (h->cand.meta != nullptr) ? R.slot : nullptr;
        out->amax = R.amax; out->reserved0 = 0;
    }
    return BW_OK;
}

// One iteration: pick (unless the candidate kernel of the previous iteration has picked already) -> step -> record,
// restarts and the candidates of the next states (rollout_enumerate).
static int rollout_iteration(bw_handle *h, const int32_t *d_index, int random_policy, uint64_t seed, bw_transition *d_slots,
                             const bw_obs_out *obs, bool already_picked, bw_transition *d_next_slots, bool *picked_next) {
    RolloutBufs &R = h->roll;
    if (!already_picked) {
        RolloutBufs Rp = R;
        if (h->cand.meta != nullptr) Rp.bits = nullptr;      // rasters are read out of the store
        launch_rollout_pick(h->P, Rp, h->cand, d_index, random_policy, seed, h->roll_step, d_slots, h->stream);
        h->launches += 1;
    }
    launch_step(h->P, R.actions, R.has_action, h->d_out, obs ? *obs : bw_obs_out{nullptr, nullptr, nullptr, nullptr}, nullptr,
                nullptr, 0, h->smem_step, h->stream);
    h->launches += 1;
    h->roll_step++;
    return rollout_enumerate(h, d_slots, h->d_out, d_next_slots, seed, picked_next);
}

int bw_rollout_commit(bw_handle *h, const int32_t *d_index, bw_transition *d_slots, const bw_obs_out *obs) {
    if (!h || !d_index || !d_slots) return BW_ERR_INVALID;
    if (!h->roll_configured) return fail(h, BW_ERR_STATE, "bw_rollout_configure / bw_rollout_begin must be called first");
    CU(cudaSetDevice(h->cfg.device));
    return rollout_iteration(h, d_index, 0, 0, d_slots, obs, false, nullptr, nullptr);
}

int bw_rollout_random(bw_handle *h, int32_t n_steps, uint64_t seed, bw_transition *d_ring, int64_t capacity, int64_t start) {
    if (!h || !d_ring || n_steps < 0 || start < 0) return BW_ERR_INVALID;
    if (!h->roll_configured) return fail(h, BW_ERR_STATE, "bw_rollout_configure / bw_rollout_begin must be called first");
    const int64_t E = h->P.E;
    if (capacity < E || capacity % E != 0 || start % E != 0)
        return fail(h, BW_ERR_INVALID, "ring capacity and start must be multiples of num_envs");
    CU(cudaSetDevice(h->cfg.device));
    // inside the call the candidate kernel of iteration k also picks for iteration k + 1 (two launches per iteration
    // with a candidate store: step, candidates); the first iteration of a call has its own pick kernel
    bool picked = false;
    for (int32_t k = 0; k < n_steps; k++) {
        bw_transition *slots = d_ring + (start + (int64_t)k * E) % capacity;
        bw_transition *next = (k + 1 < n_steps) ? d_ring + (start + (int64_t)(k + 1) * E) % capacity : nullptr;
        bool picked_next = false;
        if (int rc = rollout_iteration(h, nullptr, 1, seed, slots, nullptr, picked, next, &picked_next)) return rc;
        picked = picked_next;
    }
    return BW_OK;
}

int bw_rollout_gather_bits(bw_handle *h, const int32_t *d_env, const int32_t *d_index, int64_t n, uint64_t *d_bits) {
    if (!h || !d_index || !d_bits || n < 0) return BW_ERR_INVALID;
    if (!h->roll_configured || h->roll.slot == nullptr)
        return fail(h, BW_ERR_STATE, "bw_rollout_configure / bw_rollout_begin must be called first");
    CU(cudaSetDevice(h->cfg.device));
    if (n > 0) {
        const RolloutBufs &R = h->roll;
        const bool stored = h->cand.meta != nullptr;
        if (!stored && R.bits == nullptr) return fail(h, BW_ERR_STATE, "bw_rollout_begin must be called first");
        launch_gather_bits(h->cand, R.slot, stored ? nullptr : R.bits, R.amax, h->P.E, d_env, d_index, n, d_bits, h->stream);
        h->launches++;
    }
    CU(cudaGetLastError());
    return BW_OK;
}

int bw_unpack_transitions(bw_handle *h, const bw_transition *d_ring, const int64_t *d_indices, int64_t n, float *d_block,
                          float *d_action, float *d_next_block, float *d_binary, float *d_next_binary, float *d_reward,
                          float *d_lin_reward, uint8_t *d_done) {
    if (!h || !d_ring || n < 0) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    if (n > 0) {
        launch_unpack_transitions(d_ring, d_indices, n, d_block, d_action, d_next_block, d_binary, d_next_binary, d_reward,
                                  d_lin_reward, d_done, h->stream);
        h->launches++;
    }
    CU(cudaGetLastError());
    return BW_OK;
}

int bw_candidate_overflow(bw_handle *h, int32_t *h_needed) {
    if (!h || !h_needed) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    CU(cudaMemcpyAsync(h_needed, h->P.cand_need, sizeof(int32_t), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemsetAsync(h->P.cand_need, 0, sizeof(int32_t), h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BW_OK;
}

int bw_expand_bits(bw_handle *h, const uint64_t *d_bits, int64_t n, float *d_img) {
    if (!h || !d_bits || !d_img || n < 0) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    if (n > 0) {
        launch_expand_bits(d_bits, n, d_img, h->stream);
        h->launches++;
    }
    CU(cudaGetLastError());
    return BW_OK;
}

int bw_select_random(bw_handle *h, const bw_action *d_cand, const uint8_t *d_valid, const int32_t *d_n_cand,
                     int32_t amax, uint64_t seed, bw_action *d_actions, int32_t *d_index) {
    if (!h || !d_cand || !d_valid || !d_n_cand || !d_actions || amax <= 0) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    launch_select_random(h->P, d_cand, d_valid, d_n_cand, amax, seed, d_actions, d_index, h->stream);
    h->launches++;
    CU(cudaGetLastError());
    return BW_OK;
}

int bw_get_state(bw_handle *h, bw_block *h_blocks, int32_t *h_n_blocks) {
    if (!h || !h_blocks || !h_n_blocks) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    const int E = h->P.E;
    std::vector<Pose> pose((size_t)E * NB);
    std::vector<uint8_t> shp((size_t)E * NB);
    std::vector<uint32_t> sm(E);
    CU(cudaMemcpyAsync(pose.data(), h->P.pose, sizeof(Pose) * pose.size(), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(shp.data(), h->P.shape_of, shp.size(), cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(sm.data(), h->P.static_mask, sizeof(uint32_t) * E, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h_n_blocks, h->P.n_blocks, sizeof(int32_t) * E, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int e = 0; e < E; e++)
        for (int i = 0; i < NB; i++) {
            bw_block &b = h_blocks[(size_t)e * NB + i];
            const Pose &p = pose[(size_t)e * NB + i];
            const bool live = i < h_n_blocks[e];
            b.x = live ? p.x : 0.0; b.z = live ? p.z : 0.0; b.c = live ? p.c : 1.0; b.s = live ? p.s : 0.0;
            b.shape = live ? shp[(size_t)e * NB + i] : -1;
            b.is_static = live ? (int32_t)((sm[e] >> i) & 1u) : 0;
        }
    return BW_OK;
}

int bw_get_raster_bits(bw_handle *h, uint64_t *h_block_bits, uint64_t *h_obstacle_bits) {
    if (!h) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    const size_t bytes = sizeof(uint64_t) * (size_t)h->P.E * IMG;
    if (h_block_bits) CU(cudaMemcpyAsync(h_block_bits, h->P.block_bits, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (h_obstacle_bits) CU(cudaMemcpyAsync(h_obstacle_bits, h->P.obst_bits, bytes, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BW_OK;
}

int bw_copy_raster_bits(bw_handle *h, uint64_t *d_block_bits, uint64_t *d_obstacle_bits) {
    if (!h) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    const size_t bytes = sizeof(uint64_t) * (size_t)h->P.E * IMG;
    if (d_block_bits) CU(cudaMemcpyAsync(d_block_bits, h->P.block_bits, bytes, cudaMemcpyDeviceToDevice, h->stream));
    if (d_obstacle_bits) CU(cudaMemcpyAsync(d_obstacle_bits, h->P.obst_bits, bytes, cudaMemcpyDeviceToDevice, h->stream));
    return BW_OK;
}

int bw_get_target_state(bw_handle *h, int8_t *h_remaining, int8_t *h_reached, int32_t *h_counts) {
    if (!h || !h_remaining || !h_reached || !h_counts) return BW_ERR_INVALID;
    CU(cudaSetDevice(h->cfg.device));
    const int E = h->P.E;
    std::vector<TaskDev> tk(E);
    CU(cudaMemcpyAsync(tk.data(), h->P.task, sizeof(TaskDev) * E, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    for (int e = 0; e < E; e++) {
        for (int i = 0; i < BW_MAX_TARGETS; i++) {
            h_remaining[e * BW_MAX_TARGETS + i] = i < tk[e].n_remaining ? tk[e].remaining[i] : (int8_t)-1;
            h_reached[e * BW_MAX_TARGETS + i] = i < tk[e].n_reached ? tk[e].reached[i] : (int8_t)-1;
        }
        h_counts[2 * e] = tk[e].n_remaining;
        h_counts[2 * e + 1] = tk[e].n_reached;
    }
    return BW_OK;
}

int bw_query_placement_host(bw_handle *h, const bw_action *h_actions, const double *xlim2, const double *ylim2,
                            bw_block *h_blocks, uint8_t *h_flags) {
    if (!h || !h_actions || !h_blocks || !h_flags) return BW_ERR_INVALID;
    if (int rc = need_shapes(h)) return rc;
    CU(cudaSetDevice(h->cfg.device));
    const int E = h->P.E;
    if (!h->d_qblocks) {
        CU(dev_alloc(h, &h->d_qblocks, E));
        CU(dev_alloc(h, &h->d_qflags, E));
    }
    const double eps = 1e-6;
    const double x0 = xlim2 ? xlim2[0] : h->P.xlim0, x1 = xlim2 ? xlim2[1] : h->P.xlim1;
    const double z0 = ylim2 ? ylim2[0] : h->P.ylim0, z1 = ylim2 ? ylim2[1] : h->P.ylim1;
    CU(cudaMemcpyAsync(h->d_actions, h_actions, sizeof(bw_action) * E, cudaMemcpyHostToDevice, h->stream));
    launch_query_placement(h->P, h->d_actions, x0 - eps, x1 + eps, z0 - eps, z1 + eps, h->d_qblocks, h->d_qflags,
                           h->stream);
    h->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h_blocks, h->d_qblocks, sizeof(bw_block) * E, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h_flags, h->d_qflags, E, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BW_OK;
}

int bw_render_blocks_host(bw_handle *h, const bw_shape_desc *h_shapes, int32_t n_shapes, const bw_block *h_blocks,
                          int32_t n_blocks, const double *xlim2, const double *ylim2, uint64_t *h_bits) {
    if (!h || !h_bits || n_blocks < 0 || (n_blocks > 0 && (!h_shapes || !h_blocks))) return BW_ERR_INVALID;
    if (n_shapes < 0 || n_shapes > BW_MAX_SHAPES) return fail(h, BW_ERR_CAPACITY, "at most %d shapes", BW_MAX_SHAPES);
    if (n_blocks > 256) return fail(h, BW_ERR_CAPACITY, "at most 256 blocks per render call");
    CU(cudaSetDevice(h->cfg.device));
    if (!h->d_rshapes) {
        CU(dev_alloc(h, &h->d_rshapes, BW_MAX_SHAPES));
        CU(dev_alloc(h, &h->d_rblocks, 256));
        CU(dev_alloc(h, &h->d_rbits, IMG));
        CU(dev_alloc(h, &h->d_rxs, IMG));
        CU(dev_alloc(h, &h->d_rys, IMG));
    }
    ShapeDev dev[BW_MAX_SHAPES];
    for (int i = 0; i < n_shapes; i++) {
        if (!shape_ok(h_shapes[i])) return fail(h, BW_ERR_INVALID, "shape %d: bad face/vertex count or mass data", i);
        shape_to_dev(h_shapes[i], dev[i]);
    }
    for (int i = 0; i < n_blocks; i++)
        if (h_blocks[i].shape < 0 || h_blocks[i].shape >= n_shapes)
            return fail(h, BW_ERR_INVALID, "block %d: shape index out of range", i);
    Params P = h->P;
    if (xlim2) { P.xlim0 = xlim2[0]; P.xlim1 = xlim2[1]; }
    if (ylim2) { P.ylim0 = ylim2[0]; P.ylim1 = ylim2[1]; }
    if (!(P.xlim1 > P.xlim0) || !(P.ylim1 > P.ylim0)) return fail(h, BW_ERR_INVALID, "empty raster window");
    P.inv_step_x = (double)(IMG - 1) / (P.xlim1 - P.xlim0);
    P.inv_step_y = (double)(IMG - 1) / (P.ylim1 - P.ylim0);
    double xs[IMG], ys[IMG];
    np_linspace(P.xlim0, P.xlim1, IMG, xs);
    np_linspace(P.ylim1, P.ylim0, IMG, ys);
    CU(cudaMemcpyAsync(h->d_rxs, xs, sizeof(xs), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(h->d_rys, ys, sizeof(ys), cudaMemcpyHostToDevice, h->stream));
    if (n_shapes > 0)
        CU(cudaMemcpyAsync(h->d_rshapes, dev, sizeof(ShapeDev) * n_shapes, cudaMemcpyHostToDevice, h->stream));
    if (n_blocks > 0)
        CU(cudaMemcpyAsync(h->d_rblocks, h_blocks, sizeof(bw_block) * n_blocks, cudaMemcpyHostToDevice, h->stream));
    P.xs = h->d_rxs; P.ys = h->d_rys;
    launch_render_blocks(P, h->d_rshapes, h->d_rblocks, n_blocks, h->d_rbits, h->stream);
    h->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h_bits, h->d_rbits, sizeof(uint64_t) * IMG, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));   // xs/ys/dev live on this stack frame
    return BW_OK;
}

int bw_contains_2d_host(bw_handle *h, const bw_shape_desc *h_shape, const bw_block *h_block, const double *h_points_xz,
                        int64_t n, uint8_t *h_inside) {
    if (!h || !h_shape || !h_points_xz || !h_inside || n < 0) return BW_ERR_INVALID;
    if (!shape_ok(*h_shape)) return fail(h, BW_ERR_INVALID, "shape: bad face/vertex count or mass data");
    if (n == 0) return BW_OK;
    CU(cudaSetDevice(h->cfg.device));
    ShapeDev sd;
    shape_to_dev(*h_shape, sd);
    Pose ps;
    ps.x = h_block ? h_block->x : 0.0; ps.z = h_block ? h_block->z : 0.0;
    ps.c = h_block ? h_block->c : 1.0; ps.s = h_block ? h_block->s : 0.0;
    double *d_pts = nullptr;
    uint8_t *d_in = nullptr;
    CU(cudaMallocAsync(reinterpret_cast<void **>(&d_pts), sizeof(double) * 2 * n, h->stream));
    CU(cudaMallocAsync(reinterpret_cast<void **>(&d_in), n, h->stream));
    CU(cudaMemcpyAsync(d_pts, h_points_xz, sizeof(double) * 2 * n, cudaMemcpyHostToDevice, h->stream));
    launch_contains_points(sd, ps, d_pts, n, d_in, h->stream);
    h->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h_inside, d_in, n, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaFreeAsync(d_pts, h->stream));
    CU(cudaFreeAsync(d_in, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BW_OK;
}

int bw_get_forces(bw_handle *h, int32_t variant, bw_interface *h_itf, int32_t *h_n_itf) {
    if (!h || !h_itf || !h_n_itf || variant < 0 || variant > 1) return BW_ERR_INVALID;
    if (int rc = need_shapes(h)) return rc;
    CU(cudaSetDevice(h->cfg.device));
    const size_t E = (size_t)h->P.E;
    if (!h->d_itf) {
        CU(dev_alloc(h, &h->d_itf, E * BW_MAX_INTERFACES));
        CU(dev_alloc(h, &h->d_nitf, E));
    }
    CU(cudaMemsetAsync(h->d_itf, 0, sizeof(bw_interface) * E * BW_MAX_INTERFACES, h->stream));
    // the state did not change since the last step: re-evaluating it reproduces the same
    // interfaces and dual iterates, this time with the read-back enabled
    launch_step(h->P, nullptr, nullptr, h->d_scratch_out, bw_obs_out{nullptr, nullptr, nullptr, nullptr}, h->d_itf, h->d_nitf,
                variant, h->smem_step, h->stream);
    h->launches++;
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h_itf, h->d_itf, sizeof(bw_interface) * E * BW_MAX_INTERFACES, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(h_n_itf, h->d_nitf, sizeof(int32_t) * E, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return BW_OK;
}

}  // extern "C"

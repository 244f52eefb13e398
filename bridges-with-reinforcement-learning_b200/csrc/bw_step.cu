// K1+K2+K3 (+ the raster update of K4): one CTA (two warps) per environment.
//
//   phase 1  block placement            create_block gym_env.py:204-216 -> align_frames_2d geometry.py:39-50
//   phase 2  contact interfaces         _reset_cra_assembly assembly_env.py:281-304 -> assembly_interfaces_numpy
//   phase 3  rigid-block equilibrium    is_stable_rbe stability.py:49-71 -> rbe_solve, twice:
//            warp 0 = supports as step() leaves them (new block frozen, gym_env.py:238-245),
//            warp 1 = last block released (stabilities_freezing gym_env.py:325-333)
//   phase 4  targets / reward / termination / block_graph  gym_env.py:11-22,141-168,224-232
//   phase 5  raster update of the new block (render_blocks_2d rendering.py:105-113) + lin_reward
//
// The equilibrium check is the cone-constrained least-squares problem
//     r* = min_{f in K} ||A f - b|| / ||b||,   K = product of 2-D friction cones,
// solved in its dual (3 unknowns per free block) by a proximal-point iteration whose
// sub-problems are solved by a semismooth Newton method.  Each
// problem lives in the shared memory of its warp; reductions are warp shuffles.
#include <cstdlib>
#include "bw_common.cuh"
#include "bw_kernels.cuh"
#include "bw_solver.cuh"
#include "bw_lp.cuh"

namespace bw {

// lexicographic body pairs (a < b) over NBODY = 17 bodies
__constant__ uint8_t c_pair_a[NBODY * (NBODY - 1) / 2];
__constant__ uint8_t c_pair_b[NBODY * (NBODY - 1) / 2];
// (I >= J) index pairs of the lower triangle of a 16 x 16 block matrix, row-major
__constant__ uint8_t c_tri_i[NB * (NB + 1) / 2];
__constant__ uint8_t c_tri_j[NB * (NB + 1) / 2];

constexpr int NPAIR = NBODY * (NBODY - 1) / 2;  // 136

void upload_step_tables() {
    uint8_t pa[NPAIR], pb[NPAIR], ti[NB * (NB + 1) / 2], tj[NB * (NB + 1) / 2];
    int p = 0;
    for (int a = 0; a < NBODY; a++)
        for (int b = a + 1; b < NBODY; b++) { pa[p] = (uint8_t)a; pb[p] = (uint8_t)b; p++; }
    p = 0;
    for (int i = 0; i < NB; i++)
        for (int j = 0; j <= i; j++) { ti[p] = (uint8_t)i; tj[p] = (uint8_t)j; p++; }
    cudaMemcpyToSymbol(c_pair_a, pa, sizeof(pa));
    cudaMemcpyToSymbol(c_pair_b, pb, sizeof(pb));
    cudaMemcpyToSymbol(c_tri_i, ti, sizeof(ti));
    cudaMemcpyToSymbol(c_tri_j, tj, sizeof(tj));
    // tuning hook (tools/ only): BW_RHO_SCHEDULE="1e2,1e4,..." overrides the proximal schedule
    if (const char *env = getenv("BW_RHO_SCHEDULE")) {
        double rho[NSCHED];
        int k = 0;
        char *end = nullptr;
        for (const char *q = env; k < NSCHED && *q; q = (*end == ',') ? end + 1 : end) {
            rho[k] = strtod(q, &end);
            if (end == q) break;
            k++;
        }
        if (k > 0) {
            for (; k < NSCHED; k++) rho[k] = rho[k - 1];
            cudaMemcpyToSymbol(c_rho, rho, sizeof(rho));
        }
    }
}

// ------------------------------------------------------------------ shared memory layout
struct Layout {
    // offsets in bytes from the start of dynamic shared memory
    int pose, body, faces, pairs, contacts_G, contacts_ab, adj, prob[2], hshared, shapes, grid, total;
    int MM, MC, HS;
};

struct ProbOff {  // offsets inside one problem block
    int y, yk, d, b, g, h, f, invd, H, typ, rowbase, freebody, firstcol, size;
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

// with_h = false: the packed matrix is not part of the problem block (the two problems of the environment share
// one, see Params::share_h); o.H is then meaningless
__host__ __device__ inline ProbOff prob_layout(int MM, int MC, int HS, bool with_h = true) {
    ProbOff o;
    int p = 0;
    o.y = p; p += MM * 8;
    o.yk = p; p += MM * 8;
    o.d = p; p += MM * 8;
    o.b = p; p += MM * 8;
    o.invd = p; p += MM * 8;
    o.g = p; p += 2 * MC * 8;
    o.h = p; p += 2 * MC * 8;
    o.f = p; p += 2 * MC * 8;
    o.H = p; p += with_h ? HS * 8 : 0;
    o.typ = p; p += align16(MC);
    o.rowbase = p; p += align16(NBODY);
    o.freebody = p; p += align16(NB);
    o.firstcol = p; p += align16(NB);
    o.size = align16(p);
    return o;
}

constexpr int BODY_DOUBLES = 8;   // comx, comz, weight, depth, xmin, xmax, zmin, zmax
constexpr int FACE_DOUBLES = 8;   // nx, nz, cx, cz, e0x, e0z, e1x, e1z

__host__ __device__ inline Layout make_layout(int max_blocks, int max_itf, int n_shapes, bool share_h = false,
                                              bool lib_in_smem = true) {
    Layout L;
    L.MM = 3 * max_blocks;
    L.MC = 2 * max_itf;
    L.HS = (L.MM + 1) * (L.MM + 2) / 2;      // packed rows 0..MM (the last one carries the right-hand side)
    int p = 0;
    L.pose = p; p += NB * (int)sizeof(Pose) + align16(NB);          // poses + shape ids
    L.body = p; p += NBODY * BODY_DOUBLES * 8;
    L.contacts_G = p; p += L.MC * 12 * 8;
    L.contacts_ab = p; p += align16(2 * L.MC);
    L.adj = p; p += align16(NBODY + 1) + align16(2 * L.MC);
    ProbOff po = prob_layout(L.MM, L.MC, L.HS, !share_h);
    L.prob[0] = p; p += po.size;
    L.prob[1] = p; p += po.size;
    L.hshared = p;
    if (share_h) p += align16(L.HS * 8);
    // faces + pair scratch are dead once the contacts are assembled: they alias the
    // H matrices (the first bytes of problem 0 are y.. vectors which are initialised later)
    L.faces = L.prob[0];
    L.pairs = L.faces + NBODY * NF * FACE_DOUBLES * 8;
    int scratch_end = L.pairs + NPAIR * 16 + align16(NPAIR * 2);
    if (scratch_end > p) p = scratch_end;
    p = align16(p);
    // block library and pixel nodes: read many times per step, kept next to the problem data (unless the launch
    // trades them for one more resident CTA per SM, Params::lib_in_smem)
    L.shapes = p; p += lib_in_smem ? align16(n_shapes * (int)sizeof(ShapeDev)) : 0;
    L.grid = p; p += lib_in_smem ? 2 * IMG * 8 : 0;
    L.total = align16(p);
    return L;
}

int step_smem_bytes(int max_blocks, int max_itf, int n_shapes, bool share_h, bool lib_in_smem) {
    return make_layout(max_blocks, max_itf, n_shapes, share_h, lib_in_smem).total;
}

int step_problem_bytes(int max_blocks, int max_itf, bool share_h) {
    const Layout L = make_layout(max_blocks, max_itf, 1, share_h, true);
    const ProbOff po = prob_layout(L.MM, L.MC, L.HS, !share_h);
    return 2 * po.size + (share_h ? align16(L.HS * 8) : 0);
}
int lp_bytes(int max_blocks, int max_itf) { return lp_region_bytes(3 * max_blocks, 2 * max_itf, false); }

// The optional image outputs of a step (bw_obs_out): the raster of all blocks as f32 [1,64,64] / u8 [64,64] /
// bit-packed [64], from the CTA's shared copy `bits` (64 threads, 16-byte coalesced streaming stores).
__device__ __forceinline__ void write_obs_images(const bw_obs_out &obs, const uint64_t *bits, int e, int tid) {
    if (obs.block_img_f32 != nullptr) {
        float4 *dst = reinterpret_cast<float4 *>(obs.block_img_f32 + (size_t)e * IMG * IMG);
#pragma unroll 4
        for (int i = 0; i < IMG * IMG / 4 / 64; i++) {
            const int q = i * 64 + tid;            // float4 index: row = q / 16, nibble = q % 16
            const unsigned nib = (unsigned)(bits[q >> 4] >> (4 * (q & 15))) & 0xfu;
            __stcs(dst + q, make_float4((nib & 1u) ? 1.0f : 0.0f, (nib & 2u) ? 1.0f : 0.0f,
                                        (nib & 4u) ? 1.0f : 0.0f, (nib & 8u) ? 1.0f : 0.0f));
        }
    }
    if (obs.block_bits != nullptr) obs.block_bits[(size_t)e * IMG + tid] = bits[tid];
    if (obs.block_img_u8 != nullptr) {
        uint4 *dst = reinterpret_cast<uint4 *>(obs.block_img_u8 + (size_t)e * IMG * IMG);
#pragma unroll
        for (int i = 0; i < IMG * IMG / 16 / 64; i++) {
            const int q = i * 64 + tid;            // uint4 index: row = q / 4, 16-pixel segment = q % 4
            const unsigned seg = (unsigned)(bits[q >> 2] >> (16 * (q & 3))) & 0xffffu;
            uint4 v;
            v.x = (seg & 1u) | ((seg & 2u) << 7) | ((seg & 4u) << 14) | ((seg & 8u) << 21);
            v.y = ((seg >> 4) & 1u) | (((seg >> 4) & 2u) << 7) | (((seg >> 4) & 4u) << 14) | (((seg >> 4) & 8u) << 21);
            v.z = ((seg >> 8) & 1u) | (((seg >> 8) & 2u) << 7) | (((seg >> 8) & 4u) << 14) | (((seg >> 8) & 8u) << 21);
            v.w = ((seg >> 12) & 1u) | (((seg >> 12) & 2u) << 7) | (((seg >> 12) & 4u) << 14) | (((seg >> 12) & 8u) << 21);
            __stcs(dst + q, v);
        }
    }
}

// the environment joins the queue of the next launch (exactly once per launch, on every exit path)
__device__ __forceinline__ void enqueue_next(const Params &P, int e, int n_blocks, bool frozen_solve_needed) {
    if (!P.order_on) return;
    const int key = min(ORDER_KEYS - 1, 2 * n_blocks + (frozen_solve_needed ? 1 : 0));
    const int nxt = (P.order_phase + 1) % 3;
    const int pos = atomicAdd(&P.order_cnt[nxt * ORDER_KEYS + key], 1);
    P.order_q[((size_t)nxt * ORDER_KEYS + key) * P.E + pos] = e;
}

// statistics of the LP path (only when the handle asked for them: bw_debug_lp_stats)
//   [0..1] runs (frozen, released)  [2..3] feasible  [4..5] infeasible  [6..7] not certified
//   [8 + why] reasons of "not certified": 1 pivot cap, 2 no pivot row, 3 primal residual, 4 / 5 dual certificate, 6 setup
//   [16] pivots, [17] largest pivot count, [20] rows of all runs, [23] pivots without progress (theta = 0),
//   [24..27] runs of >= 20 pivots: count, pivots, pivots without progress, rows; [28..29] of them frozen / released,
//   [30..31] of them feasible / not
__device__ __forceinline__ void lp_count(const Params &P, const Lp &lp, int res, int which) {
    if (P.lp_stats == nullptr || lp.lane != 0) return;
    atomicAdd(&P.lp_stats[which], 1ull);
    atomicAdd(&P.lp_stats[(res == LP_FEASIBLE ? 2 : (res == LP_INFEASIBLE ? 4 : 6)) + which], 1ull);
    if (res == LP_NONE) atomicAdd(&P.lp_stats[8 + (lp.why & 7)], 1ull);
    atomicAdd(&P.lp_stats[16], (unsigned long long)lp.pivots);
    atomicMax(&P.lp_stats[17], (unsigned long long)lp.pivots);
    atomicAdd(&P.lp_stats[20], (unsigned long long)lp.m);
    atomicAdd(&P.lp_stats[23], (unsigned long long)lp.ndegen);
    if (lp.pivots >= 20) {      // the long runs: how many of them, their pivots, of which without progress, their rows
        atomicAdd(&P.lp_stats[24], 1ull);
        atomicAdd(&P.lp_stats[25], (unsigned long long)lp.pivots);
        atomicAdd(&P.lp_stats[26], (unsigned long long)lp.ndegen);
        atomicAdd(&P.lp_stats[27], (unsigned long long)lp.m);
        atomicAdd(&P.lp_stats[28 + (which & 1)], 1ull);
        atomicAdd(&P.lp_stats[30 + (res == LP_FEASIBLE ? 0 : 1)], 1ull);
    }
}

// ------------------------------------------------------------------ the kernel
// EVAL: no action is read and nothing is placed (every action counts as Action.shape = -1) -- the verdicts, distances
// and observations of the assemblies as they stand (bw_evaluate, the sweep).  Placement, raster update, the LP path
// and the target book-keeping are compiled out: a smaller image for the launches that only evaluate.
template <bool TWO, bool EVAL>
__global__ void __launch_bounds__(64, 8)
step_kernel(Params PG, const bw_action *__restrict__ actions, const uint8_t *__restrict__ mask,
            bw_step_out *__restrict__ out, bw_obs_out obs, bw_interface *__restrict__ save_itf,
            int32_t *__restrict__ save_nitf, int save_variant) {
    float *__restrict__ block_img = obs.block_img_f32;
    float *__restrict__ binary = obs.binary;
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int e = blockIdx.x;
    if (PG.order_on) {
        // heaviest environments first: ticket blockIdx.x into the queue filled by the previous launch
        __shared__ int sh_env;
        if (warp == 0) {
            const int c = PG.order_cnt[PG.order_phase * ORDER_KEYS + (ORDER_KEYS - 1 - lane)];   // lane 0 = heaviest class
            int inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += v;
            }
            const unsigned ahead = __ballot_sync(FULL, inc > (int)blockIdx.x);
            const int cls = ahead ? __ffs(ahead) - 1 : 0;
            const int before = __shfl_sync(FULL, inc - c, cls);
            if (lane == 0)
                sh_env = ahead ? PG.order_q[((size_t)PG.order_phase * ORDER_KEYS + (ORDER_KEYS - 1 - cls)) * PG.E + ((int)blockIdx.x - before)]
                               : (int)blockIdx.x;
            if (blockIdx.x == 0) PG.order_cnt[((PG.order_phase + 2) % 3) * ORDER_KEYS + lane] = 0;
        }
        __syncthreads();
        e = sh_env;
    }
    if (mask != nullptr && mask[e] == 0) {
        if (tid == 0) enqueue_next(PG, e, PG.n_blocks[e], !(PG.su_valid[e] != 0 && PG.last_out[e].stable_unfrozen != 0));
        return;
    }
    const Layout L = make_layout(PG.max_blocks, PG.max_itf, PG.n_shapes, PG.share_h != 0, PG.lib_in_smem != 0);
    // P = the handle's parameters with the block library and the pixel nodes redirected to the
    // shared-memory copies made below (every helper of bw_common.cuh reads them through P)
    Params P = PG;
    if (PG.lib_in_smem) {
        P.shapes = reinterpret_cast<const ShapeDev *>(smem + L.shapes);
        P.xs = reinterpret_cast<const double *>(smem + L.grid);
        P.ys = P.xs + IMG;
        const uint64_t *src = reinterpret_cast<const uint64_t *>(PG.shapes);
        uint64_t *dst = reinterpret_cast<uint64_t *>(smem + L.shapes);
        const int words = PG.n_shapes * (int)(sizeof(ShapeDev) / 8);
        for (int q = tid; q < words; q += 64) dst[q] = src[q];
        double *g = reinterpret_cast<double *>(smem + L.grid);
        g[tid] = PG.xs[tid];
        g[IMG + tid] = PG.ys[tid];
    }

    Pose *s_pose = reinterpret_cast<Pose *>(smem + L.pose);
    uint8_t *s_shape = smem + L.pose + NB * sizeof(Pose);
    double *s_body = reinterpret_cast<double *>(smem + L.body);       // [NBODY][8]
    double *s_face = reinterpret_cast<double *>(smem + L.faces);      // [NBODY*NF][8]
    double *s_pair_lohi = reinterpret_cast<double *>(smem + L.pairs); // [NPAIR][2]
    uint16_t *s_pair_faces = reinterpret_cast<uint16_t *>(smem + L.pairs + NPAIR * 16);
    double *s_G = reinterpret_cast<double *>(smem + L.contacts_G);
    uint8_t *s_ca = smem + L.contacts_ab;
    uint8_t *s_cb = s_ca + L.MC;
    uint8_t *s_adj_ptr = smem + L.adj;
    uint8_t *s_adj = s_adj_ptr + align16(NBODY + 1);

    __shared__ int sh_n, sh_error, sh_nitf, sh_cnt[3][2], sh_placed;
    __shared__ double sh_L0;
    __shared__ double sh_res[2];
    __shared__ int sh_status[2], sh_iters[2], sh_stable[2];
    __shared__ volatile int sh_verdict[2];   // published as soon as a solve ends; -1 = running
    __shared__ double sh_flops[2];
#ifdef BW_PROFILE
    __shared__ long long sh_prof_solve[2];
    __shared__ long long sh_prof_sub[2][6];
    __shared__ long long sh_prof_f[3];
    __shared__ long long sh_prof_lp[8];   // LP path: copy-in, setup, run, copy-out | duals + pricing, column + ratio, update, certificate
#endif
    __shared__ double sh_lin[2];
    __shared__ double sh_warm[2][3 * NB];   // starting points of the two solves (dual iterates of the last step)
    __shared__ unsigned sh_adjm[2][NBODY + 1];   // mechanism screen: contact adjacency of the free blocks
    __shared__ int sh_hlock;                // share_h: 1 while a solve owns the shared packed matrix
    __shared__ uint64_t sh_bits[IMG];      // raster of all blocks after this step
    __shared__ uint64_t sh_newbits[IMG];   // raster of the new block alone
    __shared__ double s_inv_nx[NF];        // 1 / n_x of the new block's posed faces (raster crossing estimate)
    __shared__ double sh_dist[BW_MAX_TARGETS];
    __shared__ int sh_coll[4];             // collision with: blocks, obstacles, floor, bounds
    __shared__ uint8_t sh_pair_itf[NPAIR]; // body pair -> interface index (0xFF: none); identity of the LP's columns
    __shared__ uint8_t sh_itf_pair[MAXITF];
    __shared__ int sh_lp_res[2];           // verdicts of the LP path (LP_NONE: the problem goes to screen + Newton)
    __shared__ double sh_lp_r[2];
    __shared__ int sh_lp_piv;
    __shared__ double sh_lp_flops;
    __shared__ int sh_lp_pivs[2];
    __shared__ double sh_lp_flop2[2];

#ifdef BW_PROFILE
    long long prof_t[8];
    prof_t[0] = clock64();
#define BW_STAMP(i) prof_t[i] = clock64()
#else
#define BW_STAMP(i)
#endif
    __shared__ TaskDev sh_task;            // loaded here, long before the bookkeeping needs it
    __shared__ uint8_t sh_occ[NB];
    {
        static_assert(sizeof(TaskDev) % 8 == 0 && NB % 4 == 0, "word-wise copies");
        const uint64_t *src = reinterpret_cast<const uint64_t *>(P.task + e);
        if (tid < (int)(sizeof(TaskDev) / 8)) reinterpret_cast<uint64_t *>(&sh_task)[tid] = src[tid];
        if (tid >= 32 && tid < 32 + NB / 4)
            reinterpret_cast<uint32_t *>(sh_occ)[tid - 32] = reinterpret_cast<const uint32_t *>(P.face_occ + (size_t)e * NB)[tid - 32];
    }
    bw_action act;
    if (EVAL) {
        act.target_block = -1; act.target_face = 0; act.shape = -1; act.face = 0;
        act.offset_x = 0.0; act.offset_y = 0.0; act.frozen = 0; act.reserved0 = 0;
    } else {
        act = actions[e];
    }
    const int n_old = P.n_blocks[e];
    // the LP path's stored basis: its header now, its rows on their way into L2 (they are read after the interfaces)
    const double mu_e = P.mu[e];
    LpMeta lp_meta0;
    lp_meta0.mask = 0u; lp_meta0.m = 0; lp_meta0.feasible = 0; lp_meta0.L0 = 0.0;
    if (!EVAL && PG.lp_on && act.shape >= 0) {
        lp_meta0 = PG.lp_meta[e];
        const char *rows = reinterpret_cast<const char *>(PG.lp_binv + (size_t)e * PG.lp_stride);
        const int bytes = (int)lp_meta0.m * lp_row_stride(3 * PG.max_blocks) * 8;
        if (tid == 0) asm volatile("prefetch.global.L2 [%0];" ::"l"(PG.lp_ids + (size_t)e * 3 * NB));
        if (tid >= 1 && tid < 4) asm volatile("prefetch.global.L2 [%0];" ::"l"(PG.lp_xb + (size_t)e * 3 * NB + (tid - 1) * 16));
        for (int off = tid * 128; off < bytes; off += 64 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(rows + off));
    }
    const uint64_t old_bits = P.block_bits[(size_t)e * IMG + tid];   // image row tid, used after the placement
    if (tid < NB) {     // all slots: the loads do not wait for the block count (slots past it hold stale poses)
        s_pose[tid] = P.pose[(size_t)e * NB + tid];
        s_shape[tid] = P.shape_of[(size_t)e * NB + tid];
    }
    if (tid == 0) { sh_error = 0; sh_placed = 0; sh_nitf = 0; sh_verdict[0] = -1; sh_verdict[1] = -1; sh_hlock = 0; }
    if (tid < 4) sh_coll[tid] = 0;
    if (tid < 2) sh_lp_res[tid] = LP_NONE;
    if (tid == 2) { sh_lp_piv = 0; sh_lp_flops = 0.0; sh_lp_pivs[0] = sh_lp_pivs[1] = 0; sh_lp_flop2[0] = sh_lp_flop2[1] = 0.0; }
#ifdef BW_PROFILE
    if (tid < 8) sh_prof_lp[tid] = 0;
#endif
    for (int q = tid; q < NPAIR; q += 64) sh_pair_itf[q] = 0xFF;
    // stabilities_freezing()[1] of the previous step (all of today's free blocks were free and in
    // equilibrium): the frozen solve of this step has the same rows plus contacts to a new support,
    // so that equilibrium still holds -- no solve needed
    const bool prev_released_ok = P.su_valid[e] != 0 && P.last_out[e].stable_unfrozen != 0;
    const double prev_released_res = P.last_out[e].residual_unfrozen;
    __syncthreads();

    // ---------------- phase 1: placement (thread 0)
    if (tid == 0) {
        int n = n_old;
        if (!EVAL && act.shape >= 0) {
            Pose np;
            const int err = place_block(P, s_pose, s_shape, n_old, act, np);
            if (err) sh_error = err;
            else {
                s_pose[n] = np;
                s_shape[n] = (uint8_t)act.shape;
                n = n_old + 1;
                sh_placed = 1;
            }
        }
        sh_n = n;
    }
    __syncthreads();
    const int n = sh_n;
    const int nbody = n + 1;
    const bool placed = !EVAL && sh_placed != 0;
    if (sh_error != 0) {
        // a refused action (invalid indices / environment full) leaves the state alone, but the caller's
        // buffers are still written -- the unchanged observation -- and the episode is flagged as over, so a
        // lock-step loop neither reads stale memory nor stalls on this environment
        if (tid == 0) {
            const bw_step_out prev = P.last_out[e];
            bw_step_out o;
            memset(&o, 0, sizeof(o));
            o.error = (uint8_t)sh_error;
            o.n_blocks = n_old;
            o.n_interfaces = prev.n_interfaces;
            o.stable = prev.stable; o.stable_unfrozen = prev.stable_unfrozen;
            o.residual = prev.residual; o.residual_unfrozen = prev.residual_unfrozen;
            o.collision = prev.collision; o.collision_block = prev.collision_block;
            o.collision_obstacle = prev.collision_obstacle; o.collision_floor = prev.collision_floor;
            o.collision_boundary = prev.collision_boundary;
            o.n_targets_reached = prev.n_targets_reached;
            for (int t = 0; t < BW_MAX_TARGETS; t++) o.distance_to_targets[t] = prev.distance_to_targets[t];
            o.terminated = 1;
            o.truncated = (uint8_t)(P.max_steps > 0 && n_old >= P.max_steps);
            out[e] = o;
            P.done[e] = 1;
            enqueue_next(P, e, 0, false);           // flagged as over: the next launch sees it after a reset
            if (binary != nullptr) {
                float *bf = binary + (size_t)e * 6;
                bf[0] = (float)prev.stable; bf[1] = (float)prev.collision; bf[2] = (float)prev.collision_block;
                bf[3] = (float)prev.collision_obstacle; bf[4] = (float)prev.collision_floor;
                bf[5] = (float)prev.collision_boundary;
            }
        }
        sh_bits[tid] = old_bits;
        __syncthreads();
#ifndef BW_PROFILE
        write_obs_images(obs, sh_bits, e, tid);
#endif
        return;
    }

    // ---------------- posed bodies and faces
    for (int bdy = tid; bdy < nbody; bdy += 64) {
        double *B = s_body + bdy * BODY_DOUBLES;
        if (bdy == 0) {
            B[0] = 0.0; B[1] = -0.05 * P.floor_halfwidth; B[2] = 0.0; B[3] = P.floor_depth;
            B[4] = -P.floor_halfwidth; B[5] = P.floor_halfwidth; B[6] = -0.1 * P.floor_halfwidth; B[7] = 0.0;
        } else {
            const Pose ps = s_pose[bdy - 1];
            const ShapeDev &sh = P.shapes[s_shape[bdy - 1]];
            double cx, cz;
            rot(ps.c, ps.s, sh.com_x, sh.com_z, cx, cz);
            B[0] = dadd(cx, ps.x);
            B[1] = dadd(cz, ps.z);
            B[2] = P.density * sh.area * sh.depth;
            B[3] = sh.depth;
            double xmin = 1e300, xmax = -1e300, zmin = 1e300, zmax = -1e300;
            for (int v = 0; v < sh.n_verts; v++) {
                double vx, vz;
                rot(ps.c, ps.s, sh.vert_x[v], sh.vert_z[v], vx, vz);
                vx = dadd(vx, ps.x);
                vz = dadd(vz, ps.z);
                xmin = fmin(xmin, vx); xmax = fmax(xmax, vx);
                zmin = fmin(zmin, vz); zmax = fmax(zmax, vz);
            }
            B[4] = xmin; B[5] = xmax; B[6] = zmin; B[7] = zmax;
        }
    }
    for (int q = tid; q < nbody * NF; q += 64) {
        const int bdy = q / NF, fc = q - bdy * NF;
        double *F = s_face + q * FACE_DOUBLES;
        if (bdy == 0) {
            if (fc == 0) {
                F[0] = 0.0; F[1] = 1.0; F[2] = 0.0; F[3] = 0.0;
                F[4] = -P.floor_halfwidth; F[5] = 0.0; F[6] = P.floor_halfwidth; F[7] = 0.0;
            }
        } else {
            const Pose ps = s_pose[bdy - 1];
            const ShapeDev &sh = P.shapes[s_shape[bdy - 1]];
            if (fc < sh.n_faces) {
                double ax, az;
                rot(ps.c, ps.s, sh.face_nx[fc], sh.face_nz[fc], F[0], F[1]);
                rot(ps.c, ps.s, sh.face_cx[fc], sh.face_cz[fc], ax, az);
                F[2] = dadd(ax, ps.x); F[3] = dadd(az, ps.z);
                rot(ps.c, ps.s, sh.end0_x[fc], sh.end0_z[fc], ax, az);
                F[4] = dadd(ax, ps.x); F[5] = dadd(az, ps.z);
                rot(ps.c, ps.s, sh.end1_x[fc], sh.end1_z[fc], ax, az);
                F[6] = dadd(ax, ps.x); F[7] = dadd(az, ps.z);
                if (placed && bdy == n) s_inv_nx[fc] = (F[0] != 0.0) ? 1.0 / F[0] : 0.0;
            }
        }
    }
    __syncthreads();

    // ---------------- collision flags of the last block (_check_collision assembly_env.py:346-391):
    // lanes of warp 0 = the other blocks, lanes of warp 1 = the obstacles, floor and bounds
    if (P.collision_mode != 0 && n >= 1) {
        const double tol = P.collision_tol;
        const double *BN = s_body + n * BODY_DOUBLES;      // xmin xmax zmin zmax at [4..7]
        const FaceView last = face_view_posed(s_face + (size_t)n * NF * FACE_DOUBLES, P.shapes[s_shape[n - 1]].n_faces);
        if (tid < n - 1) {
            const double *BO = s_body + (tid + 1) * BODY_DOUBLES;
            if (BO[4] <= BN[5] && BN[4] <= BO[5] && BO[6] <= BN[7] && BN[6] <= BO[7]) {     // disjoint boxes cannot penetrate
                const FaceView other = face_view_posed(s_face + (size_t)(tid + 1) * NF * FACE_DOUBLES,
                                                       P.shapes[s_shape[tid]].n_faces);
                if (polygons_collide(other, last, tol)) sh_coll[0] = 1;
            }
        } else if (tid >= 32 && tid < 32 + BW_MAX_OBSTACLES) {
            const TaskDev *tk = P.task + e;
            if (tid - 32 < tk->n_obstacles) {
                const FaceView obst = face_view_shape(*P.marker, tk->obstacle_xz[tid - 32][0], tk->obstacle_xz[tid - 32][1]);
                if (polygons_collide(obst, last, tol)) sh_coll[1] = 1;
            }
        } else if (tid == 32 + BW_MAX_OBSTACLES) {
            if (BN[6] < -tol) sh_coll[2] = 1;              // deepest vertex below the floor plane z = 0
            const Pose ps = s_pose[n - 1];                 // bounds test on the block position (x, 0, z)
            if (ps.x < P.bounds_lo[0] || ps.x > P.bounds_hi[0] || 0.0 < P.bounds_lo[1] || 0.0 > P.bounds_hi[1] ||
                ps.z < P.bounds_lo[2] || ps.z > P.bounds_hi[2])
                sh_coll[3] = 1;
        }
    }

    // ---------------- raster update of the new block (render_blocks_2d rendering.py:105-113), one thread
    // per image row, from the posed faces above (same arithmetic as pose_shape / raster_row)
    int i_lo = 0, i_hi = -1;
    {
        uint64_t bits = 0;
        if (placed) {
            const double *B = s_body + n * BODY_DOUBLES;
            const int j_lo = max((int)floor((B[4] - P.xlim0) * P.inv_step_x) - 1, 0);
            const int j_hi = min((int)ceil((B[5] - P.xlim0) * P.inv_step_x) + 1, IMG - 1);
            i_lo = max((int)floor((P.ylim1 - B[7]) * P.inv_step_y) - 1, 0);
            i_hi = min((int)ceil((P.ylim1 - B[6]) * P.inv_step_y) + 1, IMG - 1);
            if (tid >= i_lo && tid <= i_hi && j_hi >= j_lo) {
                const double *F = s_face + (size_t)n * NF * FACE_DOUBLES;
                bits = raster_row_posed(P, P.shapes[s_shape[n - 1]].n_faces, F, F + 1, F + 2, F + 3, s_inv_nx, j_lo, j_hi,
                                        tid, FACE_DOUBLES);
            }
            if (bits) P.block_bits[(size_t)e * IMG + tid] = old_bits | bits;
        }
        sh_newbits[tid] = bits;
        sh_bits[tid] = old_bits | bits;
    }

    BW_STAMP(1);
    // ---------------- phase 2: interfaces (one candidate per body pair, lexicographic order)
    // Thread per pair for the bounding-box reject; the surviving pairs of a warp are then examined one
    // after the other by the whole warp, lane = (face of a, face of b), so that the first interface in
    // (face a, face b) order is found with one test per lane instead of up to 36 tests in one thread.
    unsigned hitmask = 0;  // bit r = pair (r*64 + tid) has an interface
    for (int r = 0; r < 3; r++) {
        const int p = r * 64 + tid;
        bool near = false;
        if (p < NPAIR) {
            const int A = c_pair_a[p], B = c_pair_b[p];
            if (B < nbody) {
                const double *BA = s_body + A * BODY_DOUBLES, *BB = s_body + B * BODY_DOUBLES;
                const double slack = P.tmax + 1e-9;
                near = BA[4] <= BB[5] + slack && BB[4] <= BA[5] + slack && BA[6] <= BB[7] + slack && BB[6] <= BA[7] + slack;
            }
        }
        unsigned todo = __ballot_sync(FULL, near);
        unsigned found = 0;
#pragma unroll 1
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int pp = r * 64 + warp * 32 + src;                  // uniform in the warp
            const int A = c_pair_a[pp], B = c_pair_b[pp];
            const int nfa = (A == 0) ? 1 : P.shapes[s_shape[A - 1]].n_faces;
            const int nfb = P.shapes[s_shape[B - 1]].n_faces;
            const double dmin = fmin(s_body[A * BODY_DOUBLES + 3], s_body[B * BODY_DOUBLES + 3]);
#pragma unroll 1
            for (int fa0 = 0; fa0 < nfa; fa0 += 4) {                  // faces 0-3 of a, then 4-5
                const int fa = fa0 + (lane >> 3), fb = lane & 7;
                bool ok = false;
                double lo = 0.0, hi = 0.0;
                if (fa < nfa && fb < nfb) {
                    const double *FA = s_face + (A * NF + fa) * FACE_DOUBLES;
                    const double *FB = s_face + (B * NF + fb) * FACE_DOUBLES;
                    const double nx = FA[0], nz = FA[1], cx = FA[2], cz = FA[3];
                    if (dadd(dmul(nx, FB[0]), dmul(nz, FB[1])) < 0.0) {
                        const double d0 = dadd(dmul(dsub(FB[4], cx), nx), dmul(dsub(FB[5], cz), nz));
                        const double d1 = dadd(dmul(dsub(FB[6], cx), nx), dmul(dsub(FB[7], cz), nz));
                        if (!(fabs(d0) > P.tmax || fabs(d1) > P.tmax)) {
                            const double tx = nz, tz = -nx;
                            const double sa0 = dadd(dmul(dsub(FA[4], cx), tx), dmul(dsub(FA[5], cz), tz));
                            const double sa1 = dadd(dmul(dsub(FA[6], cx), tx), dmul(dsub(FA[7], cz), tz));
                            const double sb0 = dadd(dmul(dsub(FB[4], cx), tx), dmul(dsub(FB[5], cz), tz));
                            const double sb1 = dadd(dmul(dsub(FB[6], cx), tx), dmul(dsub(FB[7], cz), tz));
                            lo = fmax(fmin(sa0, sa1), fmin(sb0, sb1));
                            hi = fmin(fmax(sa0, sa1), fmax(sb0, sb1));
                            ok = dmul(dsub(hi, lo), dmin) >= P.amin;
                        }
                    }
                }
                const unsigned okmask = __ballot_sync(FULL, ok);
                if (okmask) {
                    if (lane == __ffs(okmask) - 1) {
                        s_pair_lohi[2 * pp] = lo;
                        s_pair_lohi[2 * pp + 1] = hi;
                        s_pair_faces[pp] = (uint16_t)(fa * 8 + fb);
                    }
                    found |= 1u << src;
                    break;
                }
            }
        }
        __syncwarp();
        const bool hit = (found >> lane) & 1u;
        const unsigned bal = __ballot_sync(FULL, hit);
        if (lane == 0) sh_cnt[r][warp] = __popc(bal);
        if (hit) hitmask |= (1u << r) | ((unsigned)__popc(bal & ((1u << lane) - 1)) << (8 + 8 * r));
    }
    {   // torque scale: largest shape radius among the blocks
        double rad = 0.0;
        if (tid < n) rad = P.shapes[s_shape[tid]].radius;
        rad = warp_max(rad);
        if (warp == 0 && lane == 0) sh_L0 = rad;
    }
    __syncthreads();
    const double invL0 = 1.0 / fmax(sh_L0, 1e-300);
    // sum of reward_img over the new block's pixels (lin_reward, successor_dqn.py:397-401): a warp
    // per image row, coalesced; the loads overlap the contact assembly below
    double lin = 0.0;
    if (placed) {
        const float *rw = PG.reward_img + (size_t)e * IMG * IMG;
#pragma unroll 2
        for (int r = i_lo + warp; r <= i_hi; r += 2) {
            const uint64_t bb = sh_newbits[r];
            if ((bb >> lane) & 1ull) lin += (double)rw[r * IMG + lane];
            if ((bb >> (lane + 32)) & 1ull) lin += (double)rw[r * IMG + lane + 32];
        }
    }
    int nitf = 0;
    for (int r = 0; r < 3; r++) nitf += sh_cnt[r][0] + sh_cnt[r][1];
    const bool overflow = nitf > P.max_itf;
    const int nc = overflow ? 0 : 2 * nitf;
    // contacts are written in pair order; the scratch (faces / pairs) is read here and dead afterwards
    if (!overflow) {
        int base = 0;
        for (int r = 0; r < 3; r++) {
            if (hitmask & (1u << r)) {
                const int p = r * 64 + tid;
                const int idx = base + (warp == 1 ? sh_cnt[r][0] : 0) + (int)((hitmask >> (8 + 8 * r)) & 0xff);
                const int A = c_pair_a[p], B = c_pair_b[p];
                const int fa = s_pair_faces[p] >> 3, fb = s_pair_faces[p] & 7;
                sh_pair_itf[p] = (uint8_t)idx;
                sh_itf_pair[idx] = (uint8_t)p;
                const double *FA = s_face + (A * NF + fa) * FACE_DOUBLES;
                const double nx = FA[0], nz = FA[1], cx = FA[2], cz = FA[3];
                const double tx = nz, tz = -nx;
                const double lohi[2] = {s_pair_lohi[2 * p], s_pair_lohi[2 * p + 1]};
                double pxs[2], pzs[2];
                for (int q = 0; q < 2; q++) {
                    const double px = dadd(cx, dmul(lohi[q], tx)), pz = dadd(cz, dmul(lohi[q], tz));
                    pxs[q] = px; pzs[q] = pz;
                    const int c = 2 * idx + q;
                    s_ca[c] = (uint8_t)A;
                    s_cb[c] = (uint8_t)B;
                    double *Gc = s_G + c * 12;
                    const double *BA = s_body + A * BODY_DOUBLES, *BB = s_body + B * BODY_DOUBLES;
                    double rx = px - BA[0], rz = pz - BA[1];
                    Gc[0] = -nx; Gc[1] = -nz; Gc[2] = -(rx * nz - rz * nx) * invL0;
                    Gc[3] = -tx; Gc[4] = -tz; Gc[5] = -(rx * tz - rz * tx) * invL0;
                    rx = px - BB[0]; rz = pz - BB[1];
                    Gc[6] = nx; Gc[7] = nz; Gc[8] = (rx * nz - rz * nx) * invL0;
                    Gc[9] = tx; Gc[10] = tz; Gc[11] = (rx * tz - rz * tx) * invL0;
                }
                if (save_itf != nullptr && idx < BW_MAX_INTERFACES) {
                    bw_interface &I = save_itf[(size_t)e * BW_MAX_INTERFACES + idx];
                    I.body_a = A - 1; I.body_b = B - 1; I.face_a = fa; I.face_b = fb;
                    I.nx = nx; I.nz = nz;
                    I.p0x = pxs[0]; I.p0z = pzs[0]; I.p1x = pxs[1]; I.p1z = pzs[1];
                    I.fn0 = I.ft0 = I.fn1 = I.ft1 = 0.0;
                }
            }
            base += sh_cnt[r][0] + sh_cnt[r][1];
        }
    }
    if (save_nitf != nullptr && tid == 0) save_nitf[e] = nitf;
    __syncthreads();

    // adjacency lists body -> contacts, ascending contact index (deterministic summation order).
    // Thread = contact point (two rounds above 64 points); per body one ballot per warp gives the
    // counts and the rank of every contact inside its 32-point segment.
    __shared__ uint8_t sh_seg_cnt[NBODY][4];          // [body][segment], segment = round * 2 + warp
    {
        int my_a[2], my_b[2], rank_a[2] = {0, 0}, rank_b[2] = {0, 0};
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int c = q * 64 + tid;
            my_a[q] = (c < nc) ? s_ca[c] : -1;
            my_b[q] = (c < nc) ? s_cb[c] : -1;
        }
        const int nq = (nc + 63) >> 6;
        const unsigned lt = (1u << lane) - 1;
#pragma unroll 1
        for (int X = 0; X < nbody; X++) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                if (q < nq) {
                    const bool fa = my_a[q] == X, fb = my_b[q] == X;
                    const unsigned bal = __ballot_sync(FULL, fa || fb);
                    if (lane == 0) sh_seg_cnt[X][q * 2 + warp] = (uint8_t)__popc(bal);
                    if (fa) rank_a[q] = __popc(bal & lt);
                    if (fb) rank_b[q] = __popc(bal & lt);
                } else if (lane == 0) {
                    sh_seg_cnt[X][q * 2 + warp] = 0;
                }
            }
        }
        __syncthreads();
        if (warp == 0) {                               // exclusive scan of the per-body totals
            int tot = 0;
            if (lane < nbody) tot = sh_seg_cnt[lane][0] + sh_seg_cnt[lane][1] + sh_seg_cnt[lane][2] + sh_seg_cnt[lane][3];
            int inc = tot;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(FULL, inc, o);
                if (lane >= o) inc += v;
            }
            if (lane <= NBODY) s_adj_ptr[lane] = (uint8_t)(inc - tot);    // bodies >= nbody: empty lists at the end
        }
        __syncthreads();
#pragma unroll
        for (int q = 0; q < 2; q++) {
            const int c = q * 64 + tid;
            if (c < nc) {
                const int seg = q * 2 + warp;
                int pa = s_adj_ptr[my_a[q]] + rank_a[q], pb = s_adj_ptr[my_b[q]] + rank_b[q];
                for (int sg = 0; sg < seg; sg++) { pa += sh_seg_cnt[my_a[q]][sg]; pb += sh_seg_cnt[my_b[q]][sg]; }
                s_adj[pa] = (uint8_t)c;
                s_adj[pb] = (uint8_t)(c | 0x80);
            }
        }
    }
    __syncthreads();

    BW_STAMP(2);
    uint32_t smask = P.static_mask[e];
    if (placed) {
        if (n >= 2) smask &= ~(1u << (n - 2));   // unfreeze_block(n-2), gym_env.py:235-236
        smask |= 1u << (n - 1);                  // action.frozen = True; freeze_block(n-1)
    }
    // ---------------- phase 3a: the verdicts of a real step as warm-started linear programmes (bw_lp.cuh).
    // Both problems start from the basis the released problem of the previous step ended with and work in the
    // memory of the two Newton problems (the basis inverses take most of it).  The frozen problem is decided
    // without a run when the previous released verdict implies it; when both are needed and their rows fit side
    // by side, warp 0 runs the frozen and warp 1 the released problem at the same time, else warp 0 runs them one
    // after the other (the released one only if the frozen one has an equilibrium: no frozen equilibrium => no
    // released one, and the episode ends there).
    const bool lp_try = !EVAL && PG.lp_on != 0 && placed && save_itf == nullptr && !overflow && nitf > 0 && n >= 1;
    if (lp_try) {
        const int region = 2 * prob_layout(L.MM, L.MC, L.HS, PG.share_h == 0).size + (PG.share_h ? align16(L.HS * 8) : 0);
        const LpOff lo = lp_layout(L.MM, L.MC, lp_region_bytes(L.MM, L.MC, true) <= region);
        unsigned char *lb = smem + L.prob[0];
        const LpMeta meta = lp_meta0;
        double *gB = PG.lp_binv + (size_t)e * PG.lp_stride;
        uint16_t *gI = PG.lp_ids + (size_t)e * 3 * NB;
        double *gX = PG.lp_xb + (size_t)e * 3 * NB;
        const uint32_t allmask = (n >= 32) ? 0xffffffffu : ((1u << n) - 1u);
        const uint32_t freeF = allmask & ~smask, freeR = freeF | (1u << (n - 1));
        const int MS = lp_row_stride(L.MM);
        const int mF = 3 * __popc(freeF), mR = mF + 3;
        // frozen problem: decided without a run when nothing is free or the previous released verdict implies it
        const bool need_F = freeF != 0u && !prev_released_ok;
        const int aF = (mF + 3) & ~3, aR = (mR + 3) & ~3;       // matrices are allocated in fours of rows (Lp::run update)
        const bool par = need_F && PG.lp_par != 0 && lo.binv == 2 * lo.size && lo.binv + (aF + aR) * MS * 8 <= region;
        const int set = par ? warp : 0;           // vector set / matrix this warp works on
        Lp lp;
        lp.G = s_G; lp.c_a = s_ca; lp.c_b = s_cb; lp.adj_ptr = s_adj_ptr; lp.adj = s_adj;
        {
            unsigned char *vb = lb + set * lo.size;
            lp.Binv = reinterpret_cast<double *>(lb + lo.binv) + (set ? aF * MS : 0);
            lp.xB = reinterpret_cast<double *>(vb + lo.xb);
            lp.pi = reinterpret_cast<double *>(vb + lo.pi);
            lp.w = reinterpret_cast<double *>(vb + lo.w);
            lp.b = reinterpret_cast<double *>(vb + lo.b);
            lp.ids = reinterpret_cast<uint16_t *>(vb + lo.ids);
            lp.pos = vb + lo.pos;
            lp.crow = reinterpret_cast<uint16_t *>(vb + lo.crow);
            lp.rowbase = reinterpret_cast<int8_t *>(vb + lo.rowbase);
            lp.freebody = vb + lo.freebody;
        }
        lp.MS = MS;
        lp.nc = nc;
        lp.lane = lane;
        lp.mu = mu_e;
        lp.pivots = 0;
        lp.flops = 0.0;
        const double r_exit = fmin(P.stable_tol, 1e-6);
        const double z_inf = fmax(1e-5, 7.0 * P.stable_tol);
        // norms of the two right-hand sides (weights of the free blocks): one reduction for both problems
        double nbR, nbF;
        {
            const bool fr = lane < n && ((freeR >> lane) & 1u);
            const double wl = fr ? s_body[(lane + 1) * BODY_DOUBLES + 2] : 0.0;
            const double ss = warp_sum(wl * wl);
            const double wn = s_body[n * BODY_DOUBLES + 2];
            nbR = sqrt(ss);
            nbF = sqrt(fmax(ss - wn * wn, 0.0));
        }
        // stored column identities and basic solution (rows lane, lane + 32): on their way while the rows are copied in
        uint16_t gid0 = LP_ART, gid1 = LP_ART;
        double gx0 = 0.0, gx1 = 0.0;
        if (par || warp == 0) {
            if (lane < (int)meta.m) { gid0 = gI[lane]; gx0 = gX[lane]; }
            if (lane + 32 < (int)meta.m) { gid1 = gI[lane + 32]; gx1 = gX[lane + 32]; }
        }
#ifdef BW_PROFILE
        long long lpt = clock64();
#define BW_LP_STAMP(i) if (tid == 0) { const long long now_ = clock64(); sh_prof_lp[i] += now_ - lpt; lpt = now_; }
#else
#define BW_LP_STAMP(i)
#endif
        const int first = need_F ? 0 : 1;
        const int trips = par ? 1 : 2 - first;
#pragma unroll 1
        for (int it = 0; it < trips; it++) {
            const int which = par ? warp : first + it;
            const uint32_t fm = which ? freeR : freeF;
            const bool active = par || warp == 0;
            // copy-in: every thread that will wait for this problem anyway takes part
            if (par) {
                Lp::load_rows(lp.Binv, gB, Lp::usable_rows(meta, fm) * MS, lane, 32);
                __syncwarp();
            } else {
                Lp::load_rows(lp.Binv, gB, Lp::usable_rows(meta, fm) * MS, tid, 64);
                __syncthreads();
            }
            BW_LP_STAMP(0);
            if (active) {
                int res = LP_NONE;
                double r = 0.0;
                const bool ready = lp.setup(fm, n, s_body, meta, gid0, gid1, gx0, gx1, sh_pair_itf, sh_L0, which ? nbR : nbF);
                BW_LP_STAMP(1);
                if (ready) res = lp.run(r_exit, z_inf, r);
                if (lane == 0) { sh_lp_res[which] = res; sh_lp_r[which] = r; sh_lp_pivs[which] = lp.pivots; sh_lp_flop2[which] = ready ? lp.flops : 0.0; }
                lp_count(PG, lp, res, which);
#ifdef BW_PROFILE
                if (tid == 0 && ready) for (int q = 0; q < 4; q++) sh_prof_lp[4 + q] += lp.tp[q];
#endif
            }
            __syncthreads();
            BW_LP_STAMP(2);
            if (!par && which == 0 && sh_lp_res[0] != LP_FEASIBLE) break;
        }
        const int resF = need_F ? sh_lp_res[0] : LP_FEASIBLE;
        // the final basis of the released problem goes back to HBM (all threads); it is only worth keeping when the
        // episode goes on, i.e. when the frozen problem has an equilibrium
        if (resF == LP_FEASIBLE && sh_lp_res[1] != LP_NONE) {
            const double *BR = reinterpret_cast<const double *>(lb + lo.binv) + (par ? aF * MS : 0);
            // (column identities, basic solution and header by the warp that ran the released problem: nobody reads
            // the stored ones any more)
            if (warp == (par ? 1 : 0)) lp.store(PG.lp_meta + e, gI, gX, freeR, sh_itf_pair, sh_lp_res[1] == LP_FEASIBLE, sh_L0);
            Lp::store_rows(gB, BR, mR * MS, tid, 64);
            BW_LP_STAMP(3);
        }
        __syncthreads();
        if (tid == 0) {
            if (resF == LP_INFEASIBLE) { sh_lp_res[1] = LP_INFEASIBLE; sh_lp_r[1] = nan(""); }
            // an answer the LP could not certify: both problems go to the Newton path; without a final released basis
            // the next step starts from an empty one
            const bool failed = (need_F && sh_lp_res[0] == LP_NONE) || (resF == LP_FEASIBLE && sh_lp_res[1] == LP_NONE);
            if (failed) { sh_lp_res[0] = LP_NONE; sh_lp_res[1] = LP_NONE; }
            if (failed || resF != LP_FEASIBLE) {
                LpMeta none;
                none.mask = 0u; none.m = 0; none.feasible = 0; none.L0 = 0.0;
                PG.lp_meta[e] = none;
            }
            sh_lp_piv = sh_lp_pivs[0] + sh_lp_pivs[1];
            sh_lp_flops = sh_lp_flop2[0] + sh_lp_flop2[1];
        }
        __syncthreads();
    }
    // ---------------- phase 3: two equilibrium problems, one warp each
    {
        const uint32_t vmask = (warp == 0) ? smask : (n > 0 ? (smask & ~(1u << (n - 1))) : smask);
        const ProbOff po = prob_layout(L.MM, L.MC, L.HS, PG.share_h == 0);
        unsigned char *pb = smem + L.prob[warp];
        Solver<TWO> S;
        S.G = s_G; S.c_a = s_ca; S.c_b = s_cb; S.adj_ptr = s_adj_ptr; S.adj = s_adj;
        S.y = reinterpret_cast<double *>(pb + po.y);
        S.yk = reinterpret_cast<double *>(pb + po.yk);
        S.d = reinterpret_cast<double *>(pb + po.d);
        S.b = reinterpret_cast<double *>(pb + po.b);
        S.g = reinterpret_cast<double *>(pb + po.g);
        S.h = reinterpret_cast<double *>(pb + po.h);
        S.f = reinterpret_cast<double *>(pb + po.f);
        S.invd = reinterpret_cast<double *>(pb + po.invd);
        // share_h: one packed matrix for both problems of the environment (the solves take turns, see below)
        S.L = reinterpret_cast<double *>(PG.share_h ? smem + L.hshared : pb + po.H);
        S.adjm = sh_adjm[warp];
        S.typ = pb + po.typ;
        S.rowbase = reinterpret_cast<int8_t *>(pb + po.rowbase);
        S.freebody = pb + po.freebody;
        S.firstcol = pb + po.firstcol;
        S.lane = lane;
        S.nc = nc;
        S.nitf = overflow ? 0 : nitf;
        S.flops = 0.0;
        S.mu = P.mu[e];
        S.inv_den = 1.0 / (1.0 + S.mu * S.mu);
        const bool want_forces = (save_itf != nullptr && warp == save_variant);
        S.r_exit = want_forces ? 1e-9 : fmin(P.stable_tol, 1e-6);
        S.exit_anytime = !want_forces;
        // released-block equilibrium => frozen-block equilibrium; no frozen equilibrium => no released one
        S.sibling = &sh_verdict[warp ^ 1];
        S.implied_by = want_forces ? -1 : (warp == 0 ? 1 : 0);
        // free blocks -> rows
        const bool is_free = (lane < n) && !((vmask >> lane) & 1u);
        const unsigned fb = __ballot_sync(FULL, is_free);
        const int nfree = __popc(fb);
        const int myrow = __popc(fb & ((1u << lane) - 1));
        if (lane == 0) S.rowbase[0] = -1;
        if (lane < n) S.rowbase[lane + 1] = is_free ? (int8_t)(3 * myrow) : (int8_t)-1;
        if (is_free) S.freebody[myrow] = (uint8_t)(lane + 1);
        S.nfree = nfree;
        S.m = 3 * nfree;
        double w = is_free ? s_body[(lane + 1) * BODY_DOUBLES + 2] : 0.0;
        const double nb = sqrt(warp_sum(w * w));
        __syncwarp();
        if (is_free) {
            S.b[3 * myrow] = 0.0;
            S.b[3 * myrow + 1] = w / nb;
            S.b[3 * myrow + 2] = 0.0;
        }
        __syncwarp();
        int stable, status = 0, iters = 0;
        double res = 0.0;
#ifdef BW_PROFILE
        for (int q = 0; q < 6; q++) S.acc_t[q] = 0;
        for (int q = 0; q < 3; q++) S.acc_f[q] = 0;
        S.t_screen = 0;
#endif
        // warm starts only along real steps (an evaluation with Action.shape = -1 may follow arbitrary
        // freeze / unfreeze calls) and never for the force read-back, whose iterates must not depend on history.
        // Starting point of the FROZEN solve: the dual iterate of the last feasible solve over (almost) the same
        // rows -- the previous step's released one when that was feasible (exactly these rows; the solve is then
        // skipped altogether below) else its frozen one (these rows minus the block released now).  Rows without
        // a stored value start at 0.  The values are parked in shared memory (the screen below uses the problem's
        // own arrays as scratch) before either warp can store new ones.  The RELEASED solve always starts from
        // y = 0: on the piles that stay in equilibrium it needs two Newton steps from there and more from the
        // previous iterate (measured: profiles/README.md, round 2), and the hard released systems are the ones
        // without equilibrium, which the mechanism screen decides.
        const bool keep_y = placed && save_itf == nullptr && P.warm_start != 0;
        bool warm = false;
        if (keep_y && warp == 0) {
            const bool ok0 = P.warm_ok[2 * e] != 0, ok1 = P.warm_ok[2 * e + 1] != 0;
            const int src = ok1 ? 1 : (ok0 ? 0 : -1);
            if (src >= 0) {
                if (is_free) {
                    const double *wy = P.warm_y + (((size_t)e * 2 + src) * NB + lane) * 3;
                    const double inb = 1.0 / nb;
                    sh_warm[warp][3 * myrow] = wy[0] * inb;
                    sh_warm[warp][3 * myrow + 1] = wy[1] * inb;
                    sh_warm[warp][3 * myrow + 2] = wy[2] * inb;
                }
                warm = true;
            }
        }
        bool have_y = false;                       // S.y holds the iterate of a solve that ended feasible
        __syncthreads();
        if (overflow) {
            stable = 0; status = 2; res = 1.0;
        } else if (nitf == 0) {
            stable = (nfree == 0);                // stability.py:53-56
            res = stable ? 0.0 : 1.0;
        } else if (nfree == 0) {
            stable = 1;
        } else if (warp == 0 && placed && prev_released_ok && !want_forces) {
            stable = 1; status = 4; res = prev_released_res;
            // the released solve of the previous step ran on exactly these rows: its iterate is this problem's
            if (warm) {
#pragma unroll 1
                for (int i = lane; i < S.m; i += 32) S.y[i] = sh_warm[warp][i];
            }
            have_y = warm;
        } else if (sh_lp_res[warp] != LP_NONE) {
            // certified by the LP path: a basic solution with ||b - A f|| <= stable_tol, or a Farkas vector
            stable = (sh_lp_res[warp] == LP_FEASIBLE);
            status = 6;
            res = sh_lp_r[warp];
        } else {
            bool certified = false;
#ifdef BW_PROFILE
            for (int q = 0; q < 6; q++) S.acc_t[q] = 0;
            for (int q = 0; q < 3; q++) S.acc_f[q] = 0;
            S.t_screen = 0;
#endif
            if (!want_forces && P.screen != 0) {
                BW_T0(t_s);
                certified = S.screen(s_body, invL0);
#ifdef BW_PROFILE
                S.t_screen = clock64() - t_s;
#endif
            }
            if (certified) {
                stable = 0; status = 5; res = nan("");   // rigid-mechanism certificate: no equilibrium, no solve
            } else {
                if (warm) {
#pragma unroll 1
                    for (int i = lane; i < S.m; i += 32) S.y[i] = sh_warm[warp][i];
                }
                // share_h: the two solves of the environment use one packed matrix and take turns.  The warp that
                // waits keeps an eye on its sibling's verdict: most of the time that verdict decides its own problem
                // too (released equilibrium => frozen equilibrium, no frozen equilibrium => no released one) and
                // the second solve never starts.
                bool mine = true;
                if (PG.share_h) {
                    int got = 0;
                    if (lane == 0) {
                        while (true) {
                            if (S.implied_by >= 0 && *S.sibling == S.implied_by) { got = 2; break; }
                            if (atomicCAS(&sh_hlock, 0, 1) == 0) { got = 1; break; }
                            __nanosleep(200);
                        }
                        // the sibling may have published between the check and the lock
                        if (got == 1 && S.implied_by >= 0 && *S.sibling == S.implied_by) { atomicExch(&sh_hlock, 0); got = 2; }
                    }
                    got = __shfl_sync(FULL, got, 0);
                    mine = (got == 1);
                }
                if (mine) {
                    status = S.solve(res, iters, warm);
                    if (PG.share_h) {
                        __syncwarp();
                        if (lane == 0) { __threadfence_block(); atomicExch(&sh_hlock, 0); }
                    }
                } else {
                    status = 3;
                }
                // out of stages with a residual already under the verdict threshold: converged within the margin
                if (status == 2 && res <= P.stable_tol) status = 0;
                if (status == 3) { stable = S.implied_by; res = nan(""); }
                else stable = (status != 2) && (res <= P.stable_tol);
                have_y = (status == 0);
            }
        }
        // only a DECIDED verdict may cut the sibling's solve short: a solve that ran out of stages (status 2,
        // the reference's stable=None) says nothing about the other problem (the reference evaluates the two
        // independently, gym_env.py:325-333)
        if (lane == 0 && status != 2) { sh_verdict[warp] = stable; __threadfence_block(); }
        if (keep_y) {
            // keep the iterate for the next step (per block, physical units); anything else invalidates the entry
            __syncwarp();
            if (have_y && lane < NB) {
                double *wy = P.warm_y + (((size_t)e * 2 + warp) * NB + lane) * 3;
                wy[0] = is_free ? S.y[3 * myrow] * nb : 0.0;
                wy[1] = is_free ? S.y[3 * myrow + 1] * nb : 0.0;
                wy[2] = is_free ? S.y[3 * myrow + 2] * nb : 0.0;
            }
            if (lane == 0) P.warm_ok[2 * e + warp] = have_y ? 1 : 0;
        }
#ifdef BW_PROFILE
        if (lane == 0) {
            sh_prof_solve[warp] = clock64() - prof_t[2];
            const bool ran = (nitf > 0 && nfree > 0 && !overflow) && !(warp == 0 && status == 4);
            for (int q = 0; q < 5; q++) sh_prof_sub[warp][q] = ran ? S.acc_t[q] : 0;
            sh_prof_sub[warp][5] = ran ? S.t_screen : 0;
            if (warp == 0) for (int q = 0; q < 3; q++) sh_prof_f[q] = ran ? S.acc_f[q] : 0;
        }
#endif
        if (lane == 0) {
            sh_res[warp] = res; sh_status[warp] = status; sh_iters[warp] = iters; sh_stable[warp] = stable;
            sh_flops[warp] = S.flops;
        }
        if (save_itf != nullptr && warp == save_variant && !overflow && nitf > 0 && nfree > 0) {
            // physical forces: f = P_K(A^T y) * ||weights||   (f[] holds the projection of the last residual())
            for (int c = lane; c < nc; c += 32) {
                bw_interface &I = save_itf[(size_t)e * BW_MAX_INTERFACES + (c >> 1)];
                if (c & 1) { I.fn1 = S.f[2 * c] * nb; I.ft1 = S.f[2 * c + 1] * nb; }
                else { I.fn0 = S.f[2 * c] * nb; I.ft0 = S.f[2 * c + 1] * nb; }
            }
        }
    }
    __syncthreads();

    BW_STAMP(4);
    // ---------------- phase 4: bookkeeping.  Warp 1: distance_to_targets (lane = block); thread 0:
    // targets, block graph, state write; then thread 0 composes the step record.
    TaskDev *task = P.task + e;
    const int stable_frozen = sh_stable[0], stable_unfrozen = sh_stable[1];
    lin = warp_sum(lin);
    if (lane == 0) sh_lin[warp] = lin;
    // (thread 0 below only touches the remaining / reached lists of sh_task, warp 1 only reads the targets)
    if (warp == 1) {
        // distance_to_targets (gym_env.py:154-160, geometry.py:89-105): min over blocks of the
        // point-to-bounding-box distance
        const int nt = sh_task.n_targets;
        for (int t = 0; t < nt; t++) {
            double dist = INFINITY;
            if (lane < n) {
                const double px = sh_task.target_xz[t][0], pz = sh_task.target_xz[t][1];
                const double *B = s_body + (lane + 1) * BODY_DOUBLES;
                const double tol = 1e-6;
                if (dsub(B[4], tol) <= px && px <= dadd(B[5], tol) && dsub(B[6], tol) <= pz && pz <= dadd(B[7], tol)) {
                    dist = 0.0;
                } else {
                    const double qx = fmin(fmax(px, B[4]), B[5]), qz = fmin(fmax(pz, B[6]), B[7]);
                    const double dx = dsub(px, qx), dz = dsub(pz, qz);
                    dist = sqrt(dadd(dmul(dx, dx), dmul(dz, dz)));
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) dist = fmin(dist, __shfl_xor_sync(FULL, dist, o));
            if (lane == 0) sh_dist[t] = dist;
        }
    } else if (tid == 0 && placed) {
        TaskDev &tk = sh_task;
        // _update_targets (gym_env.py:162-168): AABB test, removal while iterating
        const double *B = s_body + n * BODY_DOUBLES;
        const double tol = 1e-6;
        int idx = 0;
        while (idx < tk.n_remaining) {
            const int t = tk.remaining[idx];
            idx++;
            const double px = tk.target_xz[t][0], pz = tk.target_xz[t][1];
            if (dsub(B[4], tol) <= px && px <= dadd(B[5], tol) && dsub(B[6], tol) <= pz && pz <= dadd(B[7], tol) &&
                -0.5 * B[3] - tol <= 0.0 && 0.0 <= 0.5 * B[3] + tol) {
                tk.reached[tk.n_reached++] = (int8_t)t;
                // list.remove(target): first entry with equal coordinates
                int k = 0;
                for (; k < tk.n_remaining; k++) {
                    const int u = tk.remaining[k];
                    if (tk.target_xz[u][0] == px && tk.target_xz[u][1] == pz) break;
                }
                for (int q = k; q + 1 < tk.n_remaining; q++) tk.remaining[q] = tk.remaining[q + 1];
                tk.n_remaining--;
            }
        }
        // block_graph occupancy (gym_env.py:224-232)
        uint8_t *occ = P.face_occ + (size_t)e * NB;
        if (act.target_block >= 0) occ[act.target_block] = sh_occ[act.target_block] | (uint8_t)(1u << act.target_face);
        occ[n - 1] = (uint8_t)(1u << act.face);
        // the mutable tail of the task record: remaining[], reached[], n_remaining, n_reached (+ padding)
        static_assert(offsetof(TaskDev, remaining) % 8 == 0 && sizeof(TaskDev) - offsetof(TaskDev, remaining) == 16, "task tail");
        {
            const uint64_t *tail = reinterpret_cast<const uint64_t *>(&tk.remaining[0]);
            uint64_t *dst = reinterpret_cast<uint64_t *>(&task->remaining[0]);
            dst[0] = tail[0];
            dst[1] = tail[1];
        }
        P.n_blocks[e] = n;
        P.pose[(size_t)e * NB + (n - 1)] = s_pose[n - 1];
        P.shape_of[(size_t)e * NB + (n - 1)] = s_shape[n - 1];
        P.static_mask[e] = smask;
    }
    __syncthreads();
    BW_STAMP(5);
    if (tid == 0) {
        bw_step_out o;
        memset(&o, 0, sizeof(o));
        const int n_reached = sh_task.n_reached;
        for (int t = 0; t < sh_task.n_targets; t++) o.distance_to_targets[t] = sh_dist[t];
        o.stable = (uint8_t)stable_frozen;
        o.stable_unfrozen = (uint8_t)stable_unfrozen;
        o.solver_status = (uint8_t)((sh_status[0] == 2 ? 1 : 0) | (sh_status[1] == 2 ? 2 : 0) |
                                    ((sh_status[0] >= 3 && sh_status[0] != 6) ? 4 : 0) |
                                    ((sh_status[1] >= 3 && sh_status[1] != 6) ? 8 : 0) |
                                    (sh_status[0] == 6 ? 16 : 0) | (sh_status[1] == 6 ? 32 : 0));
        o.lp_pivots = (uint8_t)min(255, sh_lp_piv);
        o.residual = sh_res[0];
        o.residual_unfrozen = sh_res[1];
        o.newton_iters = sh_iters[0] + sh_iters[1];
        o.solver_kflops = (int32_t)fmin(2e9, (sh_flops[0] + sh_flops[1] + sh_lp_flops) * 1e-3);
        o.n_blocks = n;
        o.n_interfaces = nitf;
        o.n_targets_reached = (uint8_t)n_reached;
        o.error = overflow ? 2 : 0;
        const bool all_reached = sh_task.n_remaining == 0;
        o.collision_block = (uint8_t)sh_coll[0];
        o.collision_obstacle = (uint8_t)sh_coll[1];
        o.collision_floor = (uint8_t)sh_coll[2];
        o.collision_boundary = (uint8_t)sh_coll[3];
        o.collision = (uint8_t)(sh_coll[0] | sh_coll[1] | sh_coll[2] | sh_coll[3]);
        o.terminated = (uint8_t)(!stable_frozen || o.collision || all_reached);     // gym_env.py:141-144
        o.truncated = (uint8_t)(P.max_steps > 0 && n >= P.max_steps);
        if (!stable_frozen || o.collision) o.reward = -1.0f;                        // gym_env.py:11-22
        else if (!all_reached) o.reward = (float)(-1 + n_reached);
        else o.reward = (float)n_reached;
        if (placed) {
            // successor_dqn.py:397-401
            const float sum = (float)(sh_lin[0] + sh_lin[1]);
            float lr = 0.0f;
            if (stable_frozen) lr = sum / 100.0f;
            if (stable_unfrozen) lr = sum;
            o.lin_reward = lr;
            P.done[e] = (uint8_t)(o.terminated | o.truncated);
        }
        out[e] = o;
        P.last_out[e] = o;
        P.su_valid[e] = 1;
        {   // expected cost of this environment in the next launch: one more block unless the episode ended; its
            // frozen solve is skipped when the released verdict of this step was "stable"
            const bool over = placed && (o.terminated || o.truncated);
            enqueue_next(P, e, over ? 0 : (placed ? n + 1 : n), !over && !stable_unfrozen);
        }
        if (binary != nullptr) {
            float *bf = binary + (size_t)e * 6;   // get_state_features, successor_dqn.py:53-60
            bf[0] = (float)stable_frozen; bf[1] = (float)o.collision; bf[2] = (float)o.collision_block;
            bf[3] = (float)o.collision_obstacle; bf[4] = (float)o.collision_floor; bf[5] = (float)o.collision_boundary;
        }
    }
    // ---------------- phase 5: the fused observation write: f32 [1,64,64] / u8 [64,64] image of all blocks
#ifdef BW_PROFILE
    __syncthreads();
    if (tid == 0) {   // cycle counts smuggled out through fields the profile run does not need
        const long long t_end = clock64();
        out[e].distance_to_targets[0] = (double)(prof_t[1] - prof_t[0]);   // load, placement, posed faces
        out[e].distance_to_targets[1] = (double)(prof_t[2] - prof_t[1]);   // interfaces, contacts, adjacency
        out[e].distance_to_targets[2] = (double)sh_prof_solve[0];          // solve, warp 0
        out[e].distance_to_targets[3] = (double)sh_prof_solve[1];          // solve, warp 1
        out[e].residual = (double)(prof_t[5] - prof_t[4]);                 // bookkeeping
        out[e].residual_unfrozen = (double)(t_end - prof_t[5]);            // raster + lin_reward
        out[e].reward = (float)(t_end - prof_t[0]);                        // total before the image write
        // sub-phases of warp 1's solve and of the bookkeeping, packed as floats into the image buffer
        if (block_img != nullptr) {
            float *dbg = block_img + (size_t)e * IMG * IMG;
            for (int q = 0; q < 6; q++) dbg[q] = (float)sh_prof_sub[1][q];
            for (int q = 0; q < 3; q++) dbg[8 + q] = 0.0f;
            for (int q = 0; q < 6; q++) dbg[16 + q] = (float)sh_prof_sub[0][q];
            dbg[22] = (float)sh_iters[0]; dbg[23] = (float)sh_iters[1];
            for (int q = 0; q < 3; q++) dbg[24 + q] = (float)sh_prof_f[q];
            for (int q = 0; q < 8; q++) dbg[32 + q] = (float)sh_prof_lp[q];
            dbg[40] = (float)sh_lp_piv;
        }
    }
#endif
#ifndef BW_PROFILE
    write_obs_images(obs, sh_bits, e, tid);
#endif
}

void launch_step(Params &P, const bw_action *d_actions, const uint8_t *d_mask, bw_step_out *d_out,
                 const bw_obs_out &obs, bw_interface *d_itf, int32_t *d_nitf, int variant, int smem_bytes,
                 cudaStream_t stream) {
    // 3 rows per free block + the right-hand side row: one row per lane up to 10 blocks.  d_actions = nullptr:
    // evaluation only (the EVAL instantiations)
    const bool two = !(3 * P.max_blocks + 1 <= 32);
    if (d_actions == nullptr) {
        if (!two) step_kernel<false, true><<<P.E, 64, smem_bytes, stream>>>(P, nullptr, d_mask, d_out, obs, d_itf, d_nitf, variant);
        else step_kernel<true, true><<<P.E, 64, smem_bytes, stream>>>(P, nullptr, d_mask, d_out, obs, d_itf, d_nitf, variant);
    } else {
        if (!two) step_kernel<false, false><<<P.E, 64, smem_bytes, stream>>>(P, d_actions, d_mask, d_out, obs, d_itf, d_nitf, variant);
        else step_kernel<true, false><<<P.E, 64, smem_bytes, stream>>>(P, d_actions, d_mask, d_out, obs, d_itf, d_nitf, variant);
    }
    P.order_phase = (P.order_phase + 1) % 3;     // the queue this launch filled is the next one's order
}

cudaError_t configure_step(int smem_bytes) {
    cudaError_t e = cudaFuncSetAttribute(step_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(step_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(step_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(step_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
}

}  // namespace bw

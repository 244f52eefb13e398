// Host-callable launchers of the bridges_b200 kernels (one per .cu file).
#pragma once
#include "bw_common.cuh"

namespace bw {

// bw_step.cu
void upload_step_tables();
int step_smem_bytes(int max_blocks, int max_itf, int n_shapes, bool share_h, bool lib_in_smem);
cudaError_t configure_step(int smem_bytes);
int step_problem_bytes(int max_blocks, int max_itf, bool share_h);   // memory of the two Newton problems
int lp_bytes(int max_blocks, int max_itf);                           // what the LP path needs of it
void launch_step(Params &P, const bw_action *d_actions, const uint8_t *d_mask, bw_step_out *d_out,
                 const bw_obs_out &obs, bw_interface *d_itf, int32_t *d_nitf, int variant, int smem_bytes,
                 cudaStream_t stream);
double measure_fp64_gflops(cudaStream_t stream);

// bw_obs.cu
void upload_obs_tables(const float *gauss, const ShapeDev *marker);
void launch_reset(const Params &P, const bw_task *d_tasks, const uint8_t *d_mask, int only_done, cudaStream_t stream);
void launch_observe(const Params &P, float *d_block_img, float *d_binary, float *d_obstacle_img, float *d_reward_img,
                    cudaStream_t stream);
void launch_expand_bits(const uint64_t *d_bits, int64_t n, float *d_img, cudaStream_t stream);

void launch_contains_points(const ShapeDev &sh, const Pose &ps, const double *d_pts, int64_t n, uint8_t *d_inside,
                            cudaStream_t stream);
void launch_render_blocks(const Params &P, const ShapeDev *d_shapes, const bw_block *d_blocks, int n_blocks,
                          uint64_t *d_bits, cudaStream_t stream);

// bw_actions.cu
void launch_query_placement(const Params &P, const bw_action *d_actions, double xl, double xh, double zl, double zh,
                            bw_block *d_blocks, uint8_t *d_flags, cudaStream_t stream);
// store of candidate placements (enumerate_store_kernel): per environment `slots` = groups x `spg` slots, one per
// (shape, face) group and ground offset / (target block, target face, offset).  Slot word:
//   bit 31 SLOT_VALID   the slot holds the placement of its candidate (bounds flag, pixel window, raster rows)
//   bit 30 SLOT_BAD     collision_on_action (gym_env.py:304-323): a vertex outside the window
//   bit 29 SLOT_OVL     the raster overlapped the block / obstacle rasters at the call whose stamp it carries
//   bits 13-28          stamp of the last call that listed the candidate (CandCache::call of the environment)
//   bits 7-12 / 0-6     first window row / window rows (0..64)
constexpr uint32_t SLOT_VALID = 0x80000000u, SLOT_BAD = 0x40000000u, SLOT_OVL = 0x20000000u;
constexpr int SLOT_STAMP_SHIFT = 13;
constexpr uint32_t SLOT_STAMP_MASK = 0xffffu;
constexpr uint32_t SLOT_GEOM = 0x1fffu | SLOT_BAD;     // what the placement alone decides
__host__ __device__ __forceinline__ int slot_rows(uint32_t m) { return (int)(m & 0x7fu); }
__host__ __device__ __forceinline__ int slot_ilo(uint32_t m) { return (int)((m >> 7) & 0x3fu); }
struct CandCache {
    uint32_t *meta = nullptr;        // [E][slots] slot words
    uint64_t *bits = nullptr;        // [E][slots][IMG] raster rows (only the window rows are meaningful)
    Pose *pose = nullptr;            // [E][NB] pose and
    uint8_t *shape = nullptr;        // [E][NB] shape of the block the slots of (block, *) were filled for
    uint64_t *seen_block = nullptr;  // [E][IMG] block raster and
    uint64_t *seen_obst = nullptr;   // [E][IMG] obstacle raster at the last call (SLOT_OVL refers to them)
    uint32_t *call = nullptr;        // [E] stamp of the last call, 1 .. 0xffff
    int32_t slots = 0, spg = 0;
};
#ifdef __CUDACC__
// row `row` of the raster kept in slot `slot` of environment e (rows outside the window are zero)
__device__ __forceinline__ uint64_t cand_store_row(const CandCache &C, int e, int slot, int row) {
    const uint32_t m = C.meta[(size_t)e * C.slots + slot];
    const int ilo = slot_ilo(m);
    if (!(m & SLOT_VALID) || row < ilo || row >= ilo + slot_rows(m)) return 0;
    return C.bits[((size_t)e * C.slots + slot) * IMG + row];
}
#endif
// d_action_bits: dense raster copies [E,amax,64] (may be null); d_slot: slot of every listed candidate [E,amax]
// (may be null; only written with a store)
void launch_enumerate(const Params &P, const double *d_ground, int n_ground, const double *d_offsets, int n_offsets,
                      int amax, bw_action *d_cand, uint8_t *d_valid, int32_t *d_n_cand, uint64_t *d_action_bits,
                      int32_t *d_slot, const CandCache &cache, cudaStream_t stream, const uint8_t *d_mask = nullptr,
                      int32_t *d_n_valid = nullptr, const struct RollFuse *fuse = nullptr);
// fuse (fused rollout, only honoured with a candidate store): what the environment's CTA does around its candidates,
// see RollFuse below
// d_out[i] = raster of candidate d_index[i] of environment d_env[i] (d_env null: environment i), read from the
// dense copies when d_dense is given, else through d_slot from the store
void launch_gather_bits(const CandCache &cache, const int32_t *d_slot, const uint64_t *d_dense, int amax, int E,
                        const int32_t *d_env, const int32_t *d_index, int64_t n, uint64_t *d_out, cudaStream_t stream);

// 16 validity flags (0/1 bytes) of a candidate list from position `base` on, as four words with one flag per byte;
// flags at or past `cnt` (stale) read as 0
__device__ __forceinline__ void load_flags16(const uint8_t *row, int base, int cnt, bool wide, uint32_t w[4]) {
    w[0] = w[1] = w[2] = w[3] = 0;
    if (base >= cnt) return;
    if (wide) {
        const uint4 v = *reinterpret_cast<const uint4 *>(row + base);
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
    } else {
        for (int i = 0; i < 16 && base + i < cnt; i++) w[i >> 2] |= (uint32_t)row[base + i] << (8 * (i & 3));
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int keep = cnt - (base + 4 * q);
        if (keep <= 0) w[q] = 0;
        else if (keep < 4) w[q] &= (1u << (8 * keep)) - 1u;
        w[q] &= 0x01010101u;
    }
}

// Index of the k-th set flag (k < number of set flags below `cnt`), found by one warp: a lane counts 16 flags per
// trip, an inclusive scan finds the lane that holds the k-th, that lane walks its 16.  All lanes return the index.
__device__ __forceinline__ int kth_valid_candidate(const uint8_t *row, int cnt, bool wide, int k, int lane) {
    int chosen = -1;
    for (int it = 0; it * 512 < cnt; it++) {
        const int base = it * 512 + lane * 16;
        uint32_t w[4];
        load_flags16(row, base, cnt, wide, w);
        const int c = __popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += v;
        }
        const int tot = __shfl_sync(0xffffffffu, inc, 31);
        if (k < tot) {
            const bool mine = k >= inc - c && k < inc;
            if (mine) {
                int kk = k - (inc - c);
                for (int i = 0; i < 16; i++) {
                    if ((w[i >> 2] >> (8 * (i & 3))) & 1u) {
                        if (kk == 0) { chosen = base + i; break; }
                        kk--;
                    }
                }
            }
            const int src = __ffs(__ballot_sync(0xffffffffu, mine)) - 1;
            chosen = __shfl_sync(0xffffffffu, chosen, src);
            break;
        }
        k -= tot;
    }
    return chosen;
}

// A finished episode starts afresh on the task it already has (AssemblyGym.reset with the same arguments,
// gym_env.py:255-289; obstacles, targets and their rasters stay): no blocks, empty raster, targets all remaining, the
// record of an empty assembly, no warm starts.  Called by all 64 threads of the environment's CTA (thread = image
// row); used by reset_kernel (bw_reset_done) and by the rollout kernels, which restart an environment in the kernel
// that finds it finished.
__device__ __forceinline__ void restart_env(const Params &P, int e, int tid) {
    if (tid == 0) {
        TaskDev &tk = P.task[e];
        for (int i = 0; i < BW_MAX_TARGETS; i++) { tk.remaining[i] = (int8_t)i; tk.reached[i] = -1; }
        tk.n_remaining = (int8_t)tk.n_targets;
        tk.n_reached = 0;
        P.static_mask[e] = 0;
        P.n_blocks[e] = 0;
        P.done[e] = 0;
        bw_step_out o;
        memset(&o, 0, sizeof(o));
        o.stable = 1;                     // empty assembly: stability.py:53-56
        o.stable_unfrozen = 1;
        for (int i = 0; i < BW_MAX_TARGETS; i++) o.distance_to_targets[i] = INFINITY;
        P.last_out[e] = o;
        P.su_valid[e] = 1;
        P.warm_ok[2 * e] = 0;
        P.warm_ok[2 * e + 1] = 0;
        if (P.lp_meta != nullptr) {       // LpMeta (16 bytes) all zero: no basis
            reinterpret_cast<unsigned long long *>(P.lp_meta)[2 * e] = 0ull;
            reinterpret_cast<unsigned long long *>(P.lp_meta)[2 * e + 1] = 0ull;
        }
    }
    if (tid < NB) P.face_occ[(size_t)e * NB + tid] = 0;
    if (tid < IMG) P.block_bits[(size_t)e * IMG + tid] = 0;
}

// bw_rollout.cu: the kernels around step / enumerate / reset of one lock-step rollout iteration
struct RolloutBufs {
    bw_action *cand = nullptr;       // [E][amax]
    uint8_t *valid = nullptr;        // [E][amax]
    int32_t *n_cand = nullptr;       // [E]
    int32_t *n_valid = nullptr;      // [E]
    uint64_t *bits = nullptr;        // [E][amax][IMG] dense raster copies: only without a candidate store
    int32_t *slot = nullptr;         // [E][amax] store slot of every listed candidate
    bw_action *actions = nullptr;    // [E] chosen actions (input of the step kernel)
    uint8_t *has_action = nullptr;   // [E] step mask
    uint8_t *stuck = nullptr;        // [E] live environment without a candidate: reset + enumerated again
    int32_t amax = 0, env_id_base = 0;
};
// The candidate kernel of a fused-rollout iteration (enumerate_store_kernel<false, true>) closes the iteration for its
// environment: [record of the step that has just run + restart of a finished episode] -> candidates of the next state
// -> done |= "no candidate left", restart + second enumeration of an environment left without one
// (rollout_finalize_kernel) -> [the built-in random policy's pick for the next iteration].
struct RollFuse {
    RolloutBufs R;
    bw_transition *slots = nullptr;        // records of the iteration being closed (null: bw_rollout_begin, no transition)
    const bw_step_out *out = nullptr;      // step results: record + restart first (null: done by rollout_record_kernel)
    bw_transition *next_slots = nullptr;   // records of the next iteration: pick right away (null: a pick kernel follows)
    uint64_t seed = 0;
    int32_t next_step = 0;
    int32_t pad = 0;
};
void launch_rollout_pick(const Params &P, const RolloutBufs &R, const CandCache &cache, const int32_t *d_index,
                         int random_policy, uint64_t seed, int32_t step, bw_transition *d_slots, cudaStream_t stream);
void launch_rollout_record(const Params &P, const RolloutBufs &R, const bw_step_out *d_out, bw_transition *d_slots,
                           cudaStream_t stream);
void launch_rollout_finalize(const Params &P, const RolloutBufs &R, bw_transition *d_slots, cudaStream_t stream);
void launch_unpack_transitions(const bw_transition *d_ring, const int64_t *d_indices, int64_t n, float *d_block,
                               float *d_action, float *d_next_block, float *d_binary, float *d_next_binary,
                               float *d_reward, float *d_lin_reward, uint8_t *d_done, cudaStream_t stream);
void launch_select_random(const Params &P, const bw_action *d_cand, const uint8_t *d_valid, const int32_t *d_n_cand,
                          int amax, uint64_t seed, bw_action *d_actions, int32_t *d_index, cudaStream_t stream);

}  // namespace bw

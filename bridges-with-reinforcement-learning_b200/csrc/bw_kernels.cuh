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
// cache of candidate placements (enumerate_kernel<true>): per environment `slots` = groups x `spg` slots, one per
// (shape, face) group and ground offset / (target block, target face, offset)
struct CandCache {
    uint32_t *meta = nullptr;   // [E][slots] SLOT_VALID | SLOT_BAD | first window row << 8 | window rows
    uint64_t *bits = nullptr;   // [E][slots][IMG] raster rows (only the window rows are meaningful)
    Pose *pose = nullptr;       // [E][NB] pose and
    uint8_t *shape = nullptr;   // [E][NB] shape of the block the slots of (block, *) were filled for
    int32_t slots = 0, spg = 0;
};
void launch_enumerate(const Params &P, const double *d_ground, int n_ground, const double *d_offsets, int n_offsets,
                      int amax, bw_action *d_cand, uint8_t *d_valid, int32_t *d_n_cand, uint64_t *d_action_bits,
                      const CandCache &cache, cudaStream_t stream, const uint8_t *d_mask = nullptr,
                      int32_t *d_n_valid = nullptr);

// bw_rollout.cu: the kernels around step / enumerate / reset of one lock-step rollout iteration
struct RolloutBufs {
    bw_action *cand = nullptr;       // [E][amax]
    uint8_t *valid = nullptr;        // [E][amax]
    int32_t *n_cand = nullptr;       // [E]
    int32_t *n_valid = nullptr;      // [E]
    uint64_t *bits = nullptr;        // [E][amax][IMG]
    bw_action *actions = nullptr;    // [E] chosen actions (input of the step kernel)
    uint8_t *has_action = nullptr;   // [E] step mask
    uint8_t *stuck = nullptr;        // [E] live environment without a candidate: reset + enumerated again
    int32_t amax = 0, env_id_base = 0;
};
void launch_rollout_pick(const Params &P, const RolloutBufs &R, const int32_t *d_index, int random_policy, uint64_t seed,
                         int32_t step, bw_transition *d_slots, cudaStream_t stream);
void launch_rollout_record(const Params &P, const RolloutBufs &R, const bw_step_out *d_out, bw_transition *d_slots,
                           cudaStream_t stream);
void launch_rollout_finalize(const Params &P, const RolloutBufs &R, bw_transition *d_slots, cudaStream_t stream);
void launch_unpack_transitions(const bw_transition *d_ring, const int64_t *d_indices, int64_t n, float *d_block,
                               float *d_action, float *d_next_block, float *d_binary, float *d_next_binary,
                               float *d_reward, float *d_lin_reward, uint8_t *d_done, cudaStream_t stream);
void launch_select_random(const Params &P, const bw_action *d_cand, const uint8_t *d_valid, const int32_t *d_n_cand,
                          int amax, uint64_t seed, bw_action *d_actions, int32_t *d_index, cudaStream_t stream);

}  // namespace bw

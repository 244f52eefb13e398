// The kernels around step / enumerate / reset of one lock-step rollout iteration (bw_rollout_* in
// include/bridges_b200.h): the batched form of rollout_episode (robotoddler/training/successor_dqn.py:365-475).
//
//   pick      policy choice -> chosen action, first half of the transition record (state raster, action raster,
//             binary features of the state before); the built-in policy draws a uniformly random valid candidate
//   record    second half after the step kernel: next-state raster, rewards, verdicts, termination; a finished
//             episode is restarted on the spot
//   finalize  after the candidates of the next state are known: done |= "no candidate left"
//             (successor_dqn.py:409-411); such environments are restarted and flagged for a second enumeration
//   unpack    packed records -> the learner's float tensors (ReplayBuffer.sample, replay_memory.py:30-40)
//
// Records are 1.6 KB (three bit-packed rasters + scalars) instead of 3 x 16 KB float images: the rollout writes
// 1.6 MB per 1024-environment iteration, nothing here is bound by anything but launch latency.
#include "bw_common.cuh"
#include "bw_kernels.cuh"
#include "bw_rollout.cuh"

namespace bw {

static_assert(sizeof(bw_transition) == 1608, "bw_transition layout");

// One CTA of 64 threads per environment (thread = image row for the raster copies).
__global__ void __launch_bounds__(64)
rollout_pick_kernel(Params P, RolloutBufs R, CandCache C, const int32_t *__restrict__ index, int random_policy,
                    uint64_t seed, int32_t step, bw_transition *__restrict__ slots) {
    const int e = blockIdx.x;
    rollout_pick_env(P, R, C, index, random_policy, seed, step, slots, e, threadIdx.x, R.n_valid[e]);
}

// After the step: second half of the record, then a finished episode starts afresh (its task is kept) -- the reset of
// rollout_episode's next call, done by the CTA that sees the episode end.
__global__ void __launch_bounds__(64)
rollout_record_kernel(Params P, RolloutBufs R, const bw_step_out *__restrict__ out, bw_transition *__restrict__ slots) {
    const int e = blockIdx.x, tid = threadIdx.x;
    rollout_record_env(P, R, out, slots, e, tid);
    // P.done: set by the step (terminated | truncated, refused action) or by the pick (nothing to place)
    const bool finished = P.done[e] != 0;
    __syncthreads();                      // every thread has read the flag the restart clears
    if (finished) restart_env(P, e, tid);
}

// After the enumeration of the next states, CTA of 64 per environment: done |= "no candidate left"
// (successor_dqn.py:409-411); such an environment is restarted here and flagged for the second enumeration.
__global__ void __launch_bounds__(64)
rollout_finalize_kernel(Params P, RolloutBufs R, bw_transition *__restrict__ slots) {
    const int e = blockIdx.x, tid = threadIdx.x;
    // finished episodes were restarted before the enumeration: their candidates are those of a fresh environment and
    // say nothing about the recorded transition
    const int nv = R.n_valid[e];
    bool stuck = false;
    if (slots != nullptr) {
        bw_transition &T = slots[e];
        const bool live = T.valid && !T.done;
        if (live) {
            if (tid == 0) T.n_next_candidates = nv;
            stuck = nv == 0;
        } else if (!T.valid && nv == 0) {
            stuck = true;
        }
        __syncthreads();                  // every thread has read T.done
        if (live && stuck && tid == 0) T.done = 1;
    } else if (nv == 0) {
        stuck = true;
    }
    // a fresh environment always has its ground candidates; one without any is not restarted again
    if (stuck && P.n_blocks[e] == 0) stuck = false;
    __syncthreads();                      // ... and P.n_blocks, which the restart clears
    if (tid == 0) R.stuck[e] = stuck ? 1 : 0;
    if (stuck) restart_env(P, e, tid);
}

void launch_rollout_pick(const Params &P, const RolloutBufs &R, const CandCache &cache, const int32_t *d_index,
                         int random_policy, uint64_t seed, int32_t step, bw_transition *d_slots, cudaStream_t stream) {
    rollout_pick_kernel<<<P.E, 64, 0, stream>>>(P, R, cache, d_index, random_policy, seed, step, d_slots);
}

void launch_rollout_record(const Params &P, const RolloutBufs &R, const bw_step_out *d_out, bw_transition *d_slots,
                           cudaStream_t stream) {
    rollout_record_kernel<<<P.E, 64, 0, stream>>>(P, R, d_out, d_slots);
}

void launch_rollout_finalize(const Params &P, const RolloutBufs &R, bw_transition *d_slots, cudaStream_t stream) {
    rollout_finalize_kernel<<<P.E, 64, 0, stream>>>(P, R, d_slots);
}

// One CTA of 256 threads per sampled record: three 16 KB images out (float4 streaming stores), scalars by thread 0.
__global__ void __launch_bounds__(256)
unpack_transitions_kernel(const bw_transition *__restrict__ ring, const int64_t *__restrict__ indices, int64_t n,
                          float *__restrict__ block, float *__restrict__ action, float *__restrict__ next_block,
                          float *__restrict__ binary, float *__restrict__ next_binary, float *__restrict__ reward,
                          float *__restrict__ lin_reward, uint8_t *__restrict__ done) {
    const int64_t i = blockIdx.x;
    if (i >= n) return;
    const bw_transition &T = ring[indices ? indices[i] : i];
    __shared__ uint64_t s_bits[3][IMG];
    const int tid = threadIdx.x;
    if (tid < IMG) s_bits[0][tid] = T.block_bits[tid];
    else if (tid < 2 * IMG) s_bits[1][tid - IMG] = T.action_bits[tid - IMG];
    else if (tid < 3 * IMG) s_bits[2][tid - 2 * IMG] = T.next_block_bits[tid - 2 * IMG];
    __syncthreads();
    float *dst[3] = {block, action, next_block};
#pragma unroll
    for (int k = 0; k < 3; k++) {
        if (dst[k] == nullptr) continue;
        float4 *d4 = reinterpret_cast<float4 *>(dst[k] + (size_t)i * IMG * IMG);
#pragma unroll
        for (int q0 = 0; q0 < IMG * IMG / 4; q0 += 256) {
            const int q = q0 + tid;                 // float4 index: row = q / 16, nibble = q % 16
            const unsigned nib = (unsigned)(s_bits[k][q >> 4] >> (4 * (q & 15))) & 0xfu;
            __stcs(d4 + q, make_float4((nib & 1u) ? 1.0f : 0.0f, (nib & 2u) ? 1.0f : 0.0f, (nib & 4u) ? 1.0f : 0.0f,
                                       (nib & 8u) ? 1.0f : 0.0f));
        }
    }
    if (tid < 6) {
        if (binary) binary[i * 6 + tid] = (float)((T.binary >> tid) & 1u);
        if (next_binary) next_binary[i * 6 + tid] = (float)((T.next_binary >> tid) & 1u);
    }
    if (tid == 0) {
        if (reward) reward[i] = T.reward;
        if (lin_reward) lin_reward[i] = T.lin_reward;
        if (done) done[i] = T.done;
    }
}

void launch_unpack_transitions(const bw_transition *d_ring, const int64_t *d_indices, int64_t n, float *d_block,
                               float *d_action, float *d_next_block, float *d_binary, float *d_next_binary,
                               float *d_reward, float *d_lin_reward, uint8_t *d_done, cudaStream_t stream) {
    if (n <= 0) return;
    unpack_transitions_kernel<<<(unsigned)n, 256, 0, stream>>>(d_ring, d_indices, n, d_block, d_action, d_next_block,
                                                               d_binary, d_next_binary, d_reward, d_lin_reward, d_done);
}

}  // namespace bw

// Per-environment pieces of one lock-step rollout iteration (bw_rollout_* in include/bridges_b200.h): device
// functions shared by the stand-alone kernels of bw_rollout.cu and by the candidate kernel that closes an iteration
// (enumerate_store_kernel<false, true>, bw_actions.cu).  Called by every thread of the environment's CTA (64 or 128
// threads; thread = image row for the raster copies).
#pragma once
#include "bw_kernels.cuh"

namespace bw {

__device__ __forceinline__ uint64_t rmix64(uint64_t x) {   // splitmix64, as select_random_kernel
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__device__ __forceinline__ uint8_t binary_bits(const bw_step_out &o) {    // get_state_features, successor_dqn.py:53-60
    return (uint8_t)((o.stable ? 1 : 0) | (o.collision ? 2 : 0) | (o.collision_block ? 4 : 0) |
                     (o.collision_obstacle ? 8 : 0) | (o.collision_floor ? 16 : 0) | (o.collision_boundary ? 32 : 0));
}

// The policy's choice (index != nullptr) or a uniformly random valid candidate -> chosen action, step mask, first half
// of the transition record (state raster, chosen candidate's raster, action, binary features of the state before).
// nvalid: valid candidates of the environment's current list.  Contains one barrier.
__device__ __forceinline__ void rollout_pick_env(const Params &P, const RolloutBufs &R, const CandCache &C,
                                                 const int32_t *__restrict__ index, int random_policy, uint64_t seed,
                                                 int32_t step, bw_transition *__restrict__ slots, int e, int tid, int nvalid) {
    const int amax = R.amax;
    const int cnt = R.n_cand[e];
    const uint8_t *vrow = R.valid + (size_t)e * amax;
    __shared__ int s_choice;
    if (tid == 0) s_choice = -1;
    __syncthreads();
    if (nvalid > 0) {
        if (random_policy) {
            // k-th valid candidate, k uniform (counter-based hash of seed, environment and iteration): warp 0 finds it
            if (tid < 32) {
                const uint64_t r = rmix64(seed ^ rmix64((uint64_t)(R.env_id_base + e) * 0x632BE59BD9B4E019ull + (uint64_t)step));
                const bool wide = (amax & 15) == 0 && (reinterpret_cast<uintptr_t>(R.valid) & 15) == 0;
                const int a = kth_valid_candidate(vrow, cnt, wide, (int)(r % (uint64_t)nvalid), tid);
                if (tid == 0) s_choice = a;
            }
        } else if (tid == 0) {
            int a = index[e];
            if (a < 0 || a >= cnt || !vrow[a]) a = -1;     // refused: treated like "no candidate"
            s_choice = a;
        }
    }
    __syncthreads();
    const int choice = s_choice;
    bw_transition &T = slots[e];
    if (choice < 0) {
        if (tid == 0) {
            bw_action a;
            a.target_block = -1; a.target_face = 0; a.shape = -1; a.face = 0; a.offset_x = 0.0; a.offset_y = 0.0;
            a.frozen = 0; a.reserved0 = 0;
            R.actions[e] = a;
            R.has_action[e] = 0;
            T.valid = 0;
            T.done = 1;
            T.env = R.env_id_base + e;
            T.step = step;
            P.done[e] = 1;              // nothing to place: the episode is over (successor_dqn.py:409-411)
        }
        return;
    }
    if (tid < IMG) {
        T.block_bits[tid] = P.block_bits[(size_t)e * IMG + tid];
        // the chosen candidate's raster: from the dense copies, or straight out of the candidate store
        if (R.bits != nullptr) {
            T.action_bits[tid] = R.bits[((size_t)e * amax + choice) * IMG + tid];
        } else {
            const int s = R.slot[(size_t)e * amax + choice];
            T.action_bits[tid] = (s >= 0) ? cand_store_row(C, e, s, tid) : 0ull;
        }
    }
    if (tid == 0) {
        const bw_action a = R.cand[(size_t)e * amax + choice];
        R.actions[e] = a;
        R.has_action[e] = 1;
        T.action = a;
        T.binary = binary_bits(P.last_out[e]);
        T.env = R.env_id_base + e;
        T.step = step;
        T.valid = 1;
        T.n_next_candidates = 0;
    }
}

// After the step: second half of the record (next-state raster, rewards, verdicts, termination).
__device__ __forceinline__ void rollout_record_env(const Params &P, const RolloutBufs &R, const bw_step_out *__restrict__ out,
                                                   bw_transition *__restrict__ slots, int e, int tid) {
    if (!R.has_action[e]) return;
    bw_transition &T = slots[e];
    if (tid < IMG) T.next_block_bits[tid] = P.block_bits[(size_t)e * IMG + tid];
    if (tid == 0) {
        const bw_step_out o = out[e];
        T.reward = o.reward;
        T.lin_reward = o.lin_reward;
        T.next_binary = binary_bits(o);
        T.terminated = o.terminated;
        T.truncated = o.truncated;
        T.stable = o.stable;
        T.stable_unfrozen = o.stable_unfrozen;
        T.done = (uint8_t)(o.terminated | o.truncated);
        if (o.error) { T.valid = 0; T.done = 1; }      // a refused action is not a transition (the step ended the episode)
    }
}

}  // namespace bw

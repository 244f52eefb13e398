// K5: candidate actions of every environment.
//
//   generate_actions      robotoddler/utils/actions.py:7-52   (enumeration order kept)
//   get_action_features   robotoddler/training/successor_dqn.py:88-94 (one raster per candidate)
//   filter_actions        robotoddler/utils/actions.py:71-82  (collision_on_action
//                         gym_env.py:304-323 + raster overlap with blocks / obstacles)
//
// One CTA (4 warps) per environment; a warp takes one candidate at a time, its lanes are
// image rows.  Candidate rasters stay bit-packed (512 B instead of 16 KB per candidate).
#include "bw_common.cuh"
#include "bw_kernels.cuh"

namespace bw {

constexpr unsigned FULLM = 0xffffffffu;
constexpr int ENUM_WARPS = 4;

__global__ void __launch_bounds__(32 * ENUM_WARPS)
enumerate_kernel(Params P, const double *__restrict__ ground, int n_ground, const double *__restrict__ offsets,
                 int n_offsets, int amax, bw_action *__restrict__ cand, uint8_t *__restrict__ valid,
                 int32_t *__restrict__ n_cand, uint64_t *__restrict__ action_bits) {
    const int e = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    __shared__ Pose s_pose[NB];
    __shared__ uint8_t s_shape[NB];
    __shared__ uint64_t s_block[IMG], s_obst[IMG];
    __shared__ uint8_t s_free_b[NB * NF], s_free_f[NB * NF];
    __shared__ uint8_t s_grp_s[BW_MAX_SHAPES * NF], s_grp_f[BW_MAX_SHAPES * NF];
    __shared__ int s_nfree, s_ngrp;

    const int n = P.n_blocks[e];
    if (tid < n) {
        s_pose[tid] = P.pose[(size_t)e * NB + tid];
        s_shape[tid] = P.shape_of[(size_t)e * NB + tid];
    }
    if (tid < IMG) {
        s_block[tid] = P.block_bits[(size_t)e * IMG + tid];
        s_obst[tid] = P.obst_bits[(size_t)e * IMG + tid];
    }
    __syncthreads();
    if (tid == 0) {
        // receiving faces of placed blocks: all faces (assembly_env.py:153), occupied ones
        // skipped (max_blocks_per_face = 1, actions.py:42-44)
        int k = 0;
        for (int j = 0; j < n; j++) {
            const int nf = P.shapes[s_shape[j]].n_faces;
            const uint8_t occ = P.face_occ[(size_t)e * NB + j];
            for (int f = 0; f < nf; f++)
                if (!((occ >> f) & 1u)) { s_free_b[k] = (uint8_t)j; s_free_f[k] = (uint8_t)f; k++; }
        }
        s_nfree = k;
        int g = 0;
        for (int s = 0; s < P.n_shapes; s++) {
            const ShapeDev &sh = P.shapes[s];
            for (int f = 0; f < sh.n_faces; f++)
                if ((sh.target_faces_mask >> f) & 1u) { s_grp_s[g] = (uint8_t)s; s_grp_f[g] = (uint8_t)f; g++; }
        }
        s_ngrp = g;
    }
    __syncthreads();
    const int per_group = n_ground + s_nfree * n_offsets;
    const int total = s_ngrp * per_group;
    const int count = min(total, amax);
    if (tid == 0) n_cand[e] = count;
    const double eps = 1e-6;
    const double xl = dsub(P.xlim0, eps), xh = dadd(P.xlim1, eps), zl = dsub(P.ylim0, eps), zh = dadd(P.ylim1, eps);

    for (int a = warp; a < amax; a += ENUM_WARPS) {
        bw_action act;
        act.target_block = -1; act.target_face = 0; act.shape = -1; act.face = 0;
        act.offset_x = 0.0; act.offset_y = 0.0; act.frozen = 0; act.reserved0 = 0;
        bool ok = false;
        uint64_t bits0 = 0, bits1 = 0;
        if (a < count) {
            const int g = a / per_group, w = a - g * per_group;
            act.shape = s_grp_s[g];
            act.face = s_grp_f[g];
            if (w < n_ground) {
                act.offset_x = ground[w];
            } else {
                const int k = (w - n_ground) / n_offsets, oi = (w - n_ground) - k * n_offsets;
                act.target_block = s_free_b[k];
                act.target_face = s_free_f[k];
                act.offset_x = offsets[oi];
            }
            Pose ps;
            const int err = place_block(P, s_pose, s_shape, n, act, ps);
            if (err == 0) {
                const ShapeDev &sh = P.shapes[act.shape];
                // collision_on_action: any vertex outside the window (gym_env.py:304-323)
                bool outside = false;
                for (int v = 0; v < sh.n_verts; v++) {
                    double vx, vz;
                    rot(ps.c, ps.s, sh.vert_x[v], sh.vert_z[v], vx, vz);
                    vx = dadd(vx, ps.x);
                    vz = dadd(vz, ps.z);
                    if (vx < xl || vx > xh || vz < zl || vz > zh || vz < -eps) outside = true;
                }
                bits0 = raster_row(P, sh, ps, lane);
                bits1 = raster_row(P, sh, ps, lane + 32);
                const bool overlap = ((bits0 & (s_block[lane] | s_obst[lane])) != 0) ||
                                     ((bits1 & (s_block[lane + 32] | s_obst[lane + 32])) != 0);
                ok = !outside && !__any_sync(FULLM, overlap);
            } else {
                // a full environment (err 2) offers no placement; keep the candidate, mark invalid
                ok = false;
            }
        }
        const size_t o = (size_t)e * amax + a;
        if (lane == 0) {
            cand[o] = act;
            valid[o] = ok ? 1 : 0;
        }
        if (action_bits != nullptr) {
            action_bits[o * IMG + lane] = bits0;
            action_bits[o * IMG + lane + 32] = bits1;
        }
    }
}

void launch_enumerate(const Params &P, const double *d_ground, int n_ground, const double *d_offsets, int n_offsets,
                      int amax, bw_action *d_cand, uint8_t *d_valid, int32_t *d_n_cand, uint64_t *d_action_bits,
                      cudaStream_t stream) {
    enumerate_kernel<<<P.E, 32 * ENUM_WARPS, 0, stream>>>(P, d_ground, n_ground, d_offsets, n_offsets, amax, d_cand,
                                                          d_valid, d_n_cand, d_action_bits);
}

// create_block + collision_on_action for one hypothetical action per env (state untouched)
__global__ void query_placement_kernel(Params P, const bw_action *__restrict__ actions, double xl, double xh, double zl,
                                       double zh, bw_block *__restrict__ blocks, uint8_t *__restrict__ flags) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P.E) return;
    const bw_action act = actions[e];
    Pose ps;
    ps.x = ps.z = ps.s = 0.0; ps.c = 1.0;
    const int err = place_block(P, P.pose + (size_t)e * NB, P.shape_of + (size_t)e * NB, P.n_blocks[e], act, ps);
    uint8_t fl = 0;
    if (err == 1) fl |= 1;
    if (err == 2) fl |= 2;
    if (err != 1) {
        if (err == 2) {   // the pose of a block that does not fit any more is still well defined
            bw_action a2 = act;
            Params P2 = P;
            P2.max_blocks = NB + 1;
            place_block(P2, P.pose + (size_t)e * NB, P.shape_of + (size_t)e * NB, P.n_blocks[e], a2, ps);
        }
        const ShapeDev &sh = P.shapes[act.shape];
        const double eps = 1e-6;
        for (int v = 0; v < sh.n_verts; v++) {
            double vx, vz;
            rot(ps.c, ps.s, sh.vert_x[v], sh.vert_z[v], vx, vz);
            vx = dadd(vx, ps.x);
            vz = dadd(vz, ps.z);
            if (vx < xl || vx > xh || vz < zl || vz > zh || vz < -eps) fl |= 4;
        }
    }
    bw_block b;
    b.x = ps.x; b.z = ps.z; b.c = ps.c; b.s = ps.s;
    b.shape = act.shape; b.is_static = 0;
    blocks[e] = b;
    flags[e] = fl;
}

void launch_query_placement(const Params &P, const bw_action *d_actions, double xl, double xh, double zl, double zh,
                            bw_block *d_blocks, uint8_t *d_flags, cudaStream_t stream) {
    query_placement_kernel<<<(P.E + 127) / 128, 128, 0, stream>>>(P, d_actions, xl, xh, zl, zh, d_blocks, d_flags);
}

// splitmix64: counter-based, reproducible on the host
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void select_random_kernel(Params P, const bw_action *__restrict__ cand, const uint8_t *__restrict__ valid,
                                     const int32_t *__restrict__ n_cand, int amax, uint64_t seed,
                                     bw_action *__restrict__ actions, int32_t *__restrict__ index) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P.E) return;
    const int cnt = n_cand[e];
    int nvalid = 0;
    for (int a = 0; a < cnt; a++) nvalid += valid[(size_t)e * amax + a];
    bw_action act;
    act.target_block = -1; act.target_face = 0; act.shape = -1; act.face = 0;
    act.offset_x = 0.0; act.offset_y = 0.0; act.frozen = 0; act.reserved0 = 0;
    int chosen = -1;
    if (nvalid > 0) {
        const uint64_t r = mix64(seed ^ mix64((uint64_t)e * 0x632BE59BD9B4E019ull + (uint64_t)P.n_blocks[e]));
        int k = (int)(r % (uint64_t)nvalid);
        for (int a = 0; a < cnt; a++) {
            if (valid[(size_t)e * amax + a]) {
                if (k == 0) { chosen = a; break; }
                k--;
            }
        }
        act = cand[(size_t)e * amax + chosen];
    }
    actions[e] = act;
    if (index != nullptr) index[e] = chosen;
}

void launch_select_random(const Params &P, const bw_action *d_cand, const uint8_t *d_valid, const int32_t *d_n_cand,
                          int amax, uint64_t seed, bw_action *d_actions, int32_t *d_index, cudaStream_t stream) {
    select_random_kernel<<<(P.E + 127) / 128, 128, 0, stream>>>(P, d_cand, d_valid, d_n_cand, amax, seed, d_actions,
                                                                d_index);
}

}  // namespace bw

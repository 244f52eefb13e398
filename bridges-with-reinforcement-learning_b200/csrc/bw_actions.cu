// K5: candidate actions of every environment.
//
//   generate_actions      robotoddler/utils/actions.py:7-52   (enumeration order kept)
//   get_action_features   robotoddler/training/successor_dqn.py:88-94 (one raster per candidate)
//   filter_actions        robotoddler/utils/actions.py:71-82  (collision_on_action
//                         gym_env.py:304-323 + raster overlap with blocks / obstacles)
//
// One CTA (4 warps) per environment.  Candidate rasters stay bit-packed (512 B instead of 16 KB per candidate) and,
// with the candidate store (enumerate_store_kernel, the default), stay in the handle's memory from call to call
// together with their placements and overlap verdicts; enumerate_kernel is the form without a store.  In the fused
// rollout the store kernel also closes the iteration for its environment (bw_rollout.cuh).
#include "bw_common.cuh"
#include "bw_kernels.cuh"
#include "bw_rollout.cuh"

namespace bw {

__device__ __forceinline__ bool same_bits(double a, double b) {
    return __double_as_longlong(a) == __double_as_longlong(b);
}

constexpr int ENUM_THREADS = 128;
constexpr int ENUM_CHUNK = 64;        // candidates posed per pass (shared-memory tables)

// ------------------------------------------------------------------------------------------------------------
// Plain kernel (no candidate store: BW_CAND_CACHE_MB=0, or no device memory left for it).  Two phases per chunk of
// candidates.  A: one THREAD per candidate does everything that is uniform for the candidate (placement, bounds
// test, posed half-planes, pixel window) -- FP64 work that must not be replicated over the lanes of a warp.
// B: one thread per (candidate, image row inside its window) evaluates the row's bit mask and the overlap with the
// block / obstacle rasters.
__global__ void __launch_bounds__(ENUM_THREADS, 8)
enumerate_kernel(Params PG, const double *__restrict__ ground, int n_ground, const double *__restrict__ offsets,
                 int n_offsets, int amax, bw_action *__restrict__ cand, uint8_t *__restrict__ valid,
                 int32_t *__restrict__ n_cand, uint64_t *__restrict__ action_bits,
                 const uint8_t *__restrict__ mask, int32_t *__restrict__ n_valid) {
    const int e = blockIdx.x;
    if (mask != nullptr && mask[e] == 0) return;
    const int tid = threadIdx.x;
    // block library and pixel nodes in shared memory; the helpers of bw_common.cuh read them through P
    __shared__ __align__(16) unsigned char s_lib[BW_MAX_SHAPES * sizeof(ShapeDev)];
    __shared__ double s_grid[2 * IMG];
    Params P = PG;
    {
        const uint64_t *src = reinterpret_cast<const uint64_t *>(PG.shapes);
        uint64_t *dst = reinterpret_cast<uint64_t *>(s_lib);
        const int words = PG.n_shapes * (int)(sizeof(ShapeDev) / 8);
        for (int q = tid; q < words; q += ENUM_THREADS) dst[q] = src[q];
        if (tid < IMG) { s_grid[tid] = PG.xs[tid]; s_grid[IMG + tid] = PG.ys[tid]; }
        P.shapes = reinterpret_cast<const ShapeDev *>(s_lib);
        P.xs = s_grid;
        P.ys = s_grid + IMG;
    }
    __shared__ Pose s_pose[NB];
    __shared__ uint8_t s_shape[NB];
    __shared__ uint64_t s_block[IMG], s_obst[IMG];
    __shared__ uint8_t s_free_b[NB * NF], s_free_f[NB * NF];
    __shared__ uint8_t s_grp_s[BW_MAX_SHAPES * NF], s_grp_f[BW_MAX_SHAPES * NF];
    __shared__ int s_nfree, s_ngrp;
    constexpr int CHUNK = ENUM_CHUNK;
    constexpr int CPL = CHUNK / 32;             // candidates per lane in the row-count scan
    __shared__ double c_nx[CHUNK][NF], c_nz[CHUNK][NF], c_cx[CHUNK][NF], c_cz[CHUNK][NF], c_inx[CHUNK][NF];
    __shared__ int8_t c_nf[CHUNK], c_jlo[CHUNK], c_jhi[CHUNK];
    __shared__ int8_t c_ilo[CHUNK], c_bad[CHUNK];
    __shared__ int c_rowstart[CHUNK + 1];
    __shared__ int c_overlap[CHUNK];
    constexpr int OWNER_CAP = 2048;             // (candidate, row) pairs of a chunk with a direct owner entry
    __shared__ uint8_t c_owner[OWNER_CAP];      // pair -> candidate of the chunk

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
    if (tid < 32) {
        // receiving faces of placed blocks: all faces (assembly_env.py:153), occupied ones skipped
        // (max_blocks_per_face = 1, actions.py:42-44).  Lane = block: count, exclusive scan, then every lane
        // lists the free faces of its block -- same (block, face) order as the reference's nested loops.
        int nf = 0;
        unsigned freem = 0;
        if (tid < n) {
            nf = P.shapes[s_shape[tid]].n_faces;
            freem = ~(unsigned)P.face_occ[(size_t)e * NB + tid] & ((1u << nf) - 1u);
        }
        const int cnt = __popc(freem);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += v;
        }
        int k = inc - cnt;
        while (freem) {
            const int f = __ffs(freem) - 1;
            freem &= freem - 1;
            s_free_b[k] = (uint8_t)tid;
            s_free_f[k] = (uint8_t)f;
            k++;
        }
        if (tid == 31) s_nfree = inc;
    } else if (tid == 32) {
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
    if (tid == 0) {
        n_cand[e] = count;
        // generate_actions (actions.py:7-52) is unbounded: a list cut to the caller's capacity is reported
        // (bw_candidate_overflow), never dropped silently
        if (total > amax) atomicMax(P.cand_need, total);
    }
    const double eps = 1e-6;
    const double xl = dsub(P.xlim0, eps), xh = dadd(P.xlim1, eps), zl = dsub(P.ylim0, eps), zh = dadd(P.ylim1, eps);

    // only the first `count` slots are meaningful (n_cand); the rest of the caller's buffers is left alone
    int base = 0, nval = 0;
    while (base < count) {
        const int nchunk = min(CHUNK, count - base);
        // ---- phase A: thread per candidate -- the action, its pose (FP64 work that is uniform for a candidate)
        if (tid < nchunk) {
            const int a = base + tid;
            const int g = a / per_group, w = a - g * per_group;
            bw_action act;
            act.target_block = -1; act.target_face = 0; act.frozen = 0; act.reserved0 = 0; act.offset_y = 0.0;
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
            cand[(size_t)e * amax + a] = act;
            int rows = 0, ilo = 0;
            Pose ps;
            const int err = place_block(P, s_pose, s_shape, n, act, ps);
            bool bad = (err != 0);     // a full environment (err 2) offers no placement: listed but invalid
            if (!bad) {
                const ShapeDev &sh = P.shapes[act.shape];
                // collision_on_action: any vertex outside the window (gym_env.py:304-323)
                for (int v = 0; v < sh.n_verts; v++) {
                    double vx, vz;
                    rot(ps.c, ps.s, sh.vert_x[v], sh.vert_z[v], vx, vz);
                    vx = dadd(vx, ps.x);
                    vz = dadd(vz, ps.z);
                    if (vx < xl || vx > xh || vz < zl || vz > zh || vz < -eps) bad = true;
                }
                PosedShape o;
                pose_shape(P, sh, ps, o);
                for (int k = 0; k < NF; k++) {
                    c_nx[tid][k] = o.nx[k]; c_nz[tid][k] = o.nz[k]; c_cx[tid][k] = o.cx[k]; c_cz[tid][k] = o.cz[k];
                    c_inx[tid][k] = o.inv_nx[k];
                }
                c_nf[tid] = (int8_t)o.n_faces;
                c_jlo[tid] = (int8_t)o.j_lo; c_jhi[tid] = (int8_t)o.j_hi;
                ilo = o.i_lo;
                if (o.j_hi >= o.j_lo && o.i_hi >= o.i_lo) rows = o.i_hi - o.i_lo + 1;
            }
            c_ilo[tid] = (int8_t)ilo;
            c_bad[tid] = bad ? 1 : 0;
            c_overlap[tid] = 0;
            c_rowstart[tid + 1] = rows;
        }
        __syncthreads();
        if (action_bits != nullptr) {   // rows outside the windows are zero
            uint64_t *dst = action_bits + ((size_t)e * amax + base) * IMG;
            for (int q = tid; q < nchunk * IMG; q += ENUM_THREADS) dst[q] = 0;
        }
        if (tid < 32) {
            // inclusive scan of the row counts (CPL candidates per lane), c_rowstart[t + 1] = rows of 0..t
            int av[CPL], sum = 0;
#pragma unroll
            for (int j = 0; j < CPL; j++) {
                const int idx = CPL * tid + j;
                av[j] = (idx < nchunk) ? c_rowstart[idx + 1] : 0;
                sum += av[j];
            }
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, inc, o);
                if (tid >= o) inc += v;
            }
            int run = inc - sum;
#pragma unroll
            for (int j = 0; j < CPL; j++) {
                const int idx = CPL * tid + j;
                // the pairs run .. run + av[j] - 1 belong to candidate idx (phase B looks its pair up here)
                for (int r = run; r < run + av[j] && r < OWNER_CAP; r++) c_owner[r] = (uint8_t)idx;
                run += av[j];
                if (idx < nchunk) c_rowstart[idx + 1] = run;
            }
            if (tid == 0) c_rowstart[0] = 0;
        }
        __syncthreads();
        // ---- phase B: thread per (candidate, row of its window)
        const int npairs = c_rowstart[nchunk];
        for (int q = tid; q < npairs; q += ENUM_THREADS) {
            int lo = 0;                             // largest t with rowstart[t] <= q
            if (q < OWNER_CAP) {
                lo = c_owner[q];
            } else {
                int hi = nchunk - 1;
                while (lo < hi) {
                    const int mid = (lo + hi + 1) >> 1;
                    if (c_rowstart[mid] <= q) lo = mid; else hi = mid - 1;
                }
            }
            const int t = lo, row = c_ilo[lo] + (q - c_rowstart[lo]);
            const uint64_t bits = raster_row_posed_mixed(P, c_nf[t], c_nx[t], c_nz[t], c_cx[t], c_cz[t], c_inx[t],
                                                         c_jlo[t], c_jhi[t], row);
            if (bits & (s_block[row] | s_obst[row])) atomicOr(&c_overlap[t], 1);
            if (action_bits != nullptr && bits) action_bits[((size_t)e * amax + base + t) * IMG + row] = bits;
        }
        __syncthreads();
        const bool keep = tid < nchunk && !c_bad[tid] && !c_overlap[tid];
        if (tid < nchunk) valid[(size_t)e * amax + base + tid] = keep ? 1 : 0;
        nval += __syncthreads_count(keep);
        base += nchunk;
    }
    if (n_valid != nullptr && tid == 0) n_valid[e] = nval;
}

// ------------------------------------------------------------------------------------------------------------
// Kernel with the candidate store (CandCache, bw_kernels.cuh).  A candidate is (group, ground offset) or (group,
// target block, target face, offset): its pose, its bounds flag and its raster depend on nothing but the block
// library and the pose of the target block, so they survive from call to call in the slot of that candidate.  A
// block whose pose or shape differs from the copy taken when its slots were filled (a reset, a new block)
// invalidates its slots at the start of the call; ground slots live until the library or the offset tables change
// (host side).  What changes from call to call is the overlap with the block / obstacle rasters -- and within an
// episode those only grow: the store keeps the rasters the last call saw (seen_block, seen_obst) and every slot
// the overlap verdict of that call (SLOT_OVL) with the call's stamp.  A candidate that was listed by the previous
// call is therefore tested against the NEW pixels only (a handful of rows, or none at all once it overlaps);
// a raster that lost pixels or whose obstacles changed (a reset) makes the call a fresh one: every listed
// candidate is tested against the whole raster again.
//
// One CTA per environment.  Per chunk of CCH candidates:
//   A0  thread per candidate: the action, its slot, hit / miss, the rows that need a test -> two dense lists
//   Bh  hits with rows to test: 8 lanes per candidate, rows read back from the store
//   A1  misses, DENSE (MISS_CAP per round): eight lanes pose the m-th miss (a vertex and a face each)
//   Bm  sixteen lanes per miss, a row of its window each: the row's bit mask (exact where it matters), stored, tested
//   C   thread per candidate: validity flag, slot word with the new verdict and stamp
// COPY: the caller wants the rasters copied out ([E,amax,64]); every listed candidate is then read and tested in
// full.  Without it the caller gets d_slot (where the raster lives) and gathers what it needs.
constexpr int CCH = 256;
constexpr int CPT = CCH / ENUM_THREADS;
constexpr int MISS_CAP = 32;
constexpr uint32_t W_INCR = 0x80000000u;          // working word only: the stored verdict holds, test new pixels only

__device__ __forceinline__ int list_append(bool pred, int *counter, int lane) {
    const unsigned m = __ballot_sync(0xffffffffu, pred);
    int start = 0;
    if (lane == 0 && m) start = atomicAdd(counter, __popc(m));
    start = __shfl_sync(0xffffffffu, start, 0);
    return start + __popc(m & ((1u << lane) - 1u));
}

// The candidates of environment e (the whole CTA; P.shapes / P.xs / P.ys point into shared memory and the first
// barrier inside also covers the caller's copy).  Returns the number of valid candidates (all threads).
template <bool COPY>
__device__ __forceinline__ int
enumerate_store_env(const Params &P, const int e, const double *__restrict__ ground, int n_ground,
                    const double *__restrict__ offsets, int n_offsets, int amax, bw_action *__restrict__ cand,
                    uint8_t *__restrict__ valid, int32_t *__restrict__ n_cand, uint64_t *__restrict__ action_bits,
                    int32_t *__restrict__ slot_out, const CandCache &C, int32_t *__restrict__ n_valid) {
    const int tid = threadIdx.x, lane = tid & 31;
    __shared__ Pose s_pose[NB];
    __shared__ uint8_t s_shape[NB];
    __shared__ uint64_t s_full[IMG], s_delta[IMG];      // blocks | obstacles; the pixels the last call had not seen
    __shared__ uint8_t s_free_b[NB * NF], s_free_f[NB * NF];
    __shared__ uint8_t s_grp_s[BW_MAX_SHAPES * NF], s_grp_f[BW_MAX_SHAPES * NF];
    __shared__ int s_nfree, s_ngrp;
    __shared__ unsigned s_inval, s_dmask[2], s_prev;
    // posed-face tables of the misses of one round
    __shared__ double t_nx[MISS_CAP][NF], t_nz[MISS_CAP][NF], t_cx[MISS_CAP][NF], t_cz[MISS_CAP][NF], t_inx[MISS_CAP][NF];
    __shared__ int8_t t_nf[MISS_CAP], t_jlo[MISS_CAP], t_jhi[MISS_CAP], t_ilo[MISS_CAP], t_rows[MISS_CAP];
    // per candidate of the chunk
    __shared__ uint32_t c_w[CCH];                       // SLOT_BAD | SLOT_OVL | window (slot word layout) | W_INCR
    __shared__ int c_slot[CCH];
    __shared__ uint8_t c_wlo[CCH], c_wn[CCH];           // rows a hit has to test
    __shared__ uint16_t s_miss[CCH], s_work[CCH];
    __shared__ int s_nmiss, s_nwork, s_nval;

    const int n = P.n_blocks[e];
    Pose my_pose, kept_pose;                  // lane = block: its pose and the one its slots were filled for
    uint8_t my_shape = 0, kept_shape = 0;
    if (tid < NB) {                           // all loads of the prologue are in flight together
        my_pose = P.pose[(size_t)e * NB + tid];
        my_shape = P.shape_of[(size_t)e * NB + tid];
        kept_pose = C.pose[(size_t)e * NB + tid];
        kept_shape = C.shape[(size_t)e * NB + tid];
    }
    if (tid < n) {
        s_pose[tid] = my_pose;
        s_shape[tid] = my_shape;
    }
    uint64_t regress = 0;
    if (tid < IMG) {
        const size_t i = (size_t)e * IMG + tid;
        const uint64_t b = P.block_bits[i], o = P.obst_bits[i];
        const uint64_t sb = C.seen_block[i], so = C.seen_obst[i];
        regress = (sb & ~b) | (so ^ o);
        s_full[tid] = b | o;
        s_delta[tid] = b & ~sb;
        C.seen_block[i] = b;
        C.seen_obst[i] = o;
    }
    if (tid == 0) {
        // stamp of this call (16 bits in every slot word it touches): 1 .. 0xffff, then the slots are wiped
        s_prev = C.call[e];
        s_nmiss = 0; s_nwork = 0; s_nval = 0; s_inval = 0;
    }
    const bool fresh = __syncthreads_or(regress != 0) != 0;
    if (tid < IMG) {
        if (fresh) s_delta[tid] = s_full[tid];
        const unsigned nz = __ballot_sync(0xffffffffu, s_delta[tid] != 0);
        if (lane == 0) s_dmask[tid >> 5] = nz;
    }
    unsigned prev = s_prev;
    const bool wipe = prev >= 0xffffu;
    const unsigned cur = wipe ? 1u : prev + 1u;
    if (wipe) prev = 0xffffffffu;                       // matches no stamp
    if (tid == 0) C.call[e] = cur;
    if (tid < 32) {
        // receiving faces of placed blocks: all faces (assembly_env.py:153), occupied ones skipped
        // (max_blocks_per_face = 1, actions.py:42-44).  Lane = block: count, exclusive scan, then every lane
        // lists the free faces of its block -- same (block, face) order as the reference's nested loops.
        int nf = 0;
        unsigned freem = 0;
        if (tid < n) {
            nf = P.shapes[s_shape[tid]].n_faces;
            freem = ~(unsigned)P.face_occ[(size_t)e * NB + tid] & ((1u << nf) - 1u);
        }
        const int cnt = __popc(freem);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (tid >= o) inc += v;
        }
        int k = inc - cnt;
        while (freem) {
            const int f = __ffs(freem) - 1;
            freem &= freem - 1;
            s_free_b[k] = (uint8_t)tid;
            s_free_f[k] = (uint8_t)f;
            k++;
        }
        if (tid == 31) s_nfree = inc;
    } else if (tid < 64) {
        // candidate groups = (shape, face) pairs with the target_faces bit, shape-major: warp 1, two pairs per lane
        static_assert(BW_MAX_SHAPES * NF <= 64, "two (shape, face) pairs per lane");
        const int l = tid - 32;
        unsigned has = 0;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            const int idx = 2 * l + j, s = idx / NF, f = idx - s * NF;
            if (s < P.n_shapes && f < P.shapes[s].n_faces && ((P.shapes[s].target_faces_mask >> f) & 1u)) has |= 1u << j;
        }
        const int cnt = __popc(has);
        int inc = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, o);
            if (l >= o) inc += v;
        }
        int g = inc - cnt;
#pragma unroll
        for (int j = 0; j < 2; j++) {
            if ((has >> j) & 1u) {
                const int idx = 2 * l + j, s = idx / NF;
                s_grp_s[g] = (uint8_t)s;
                s_grp_f[g] = (uint8_t)(idx - s * NF);
                g++;
            }
        }
        if (l == 31) s_ngrp = inc;
    }
    // blocks that are not the ones their slots were filled for: drop those slots, remember the new block
    if (tid < n) {
        const bool same = same_bits(kept_pose.x, my_pose.x) && same_bits(kept_pose.z, my_pose.z) &&
                          same_bits(kept_pose.c, my_pose.c) && same_bits(kept_pose.s, my_pose.s) && kept_shape == my_shape;
        if (!same) {
            C.pose[(size_t)e * NB + tid] = my_pose;
            C.shape[(size_t)e * NB + tid] = my_shape;
            atomicOr(&s_inval, 1u << tid);
        }
    }
    __syncthreads();
    uint32_t *meta = C.meta + (size_t)e * C.slots;
    const uint64_t *store = C.bits + (size_t)e * C.slots * IMG;
    if (wipe) {
        for (int q = tid; q < C.slots; q += ENUM_THREADS) meta[q] = 0;
    } else {
        unsigned inv = s_inval;
        const int per_block = NF * n_offsets;
        while (inv) {
            const int b = __ffs(inv) - 1;
            inv &= inv - 1;
            for (int q = tid; q < s_ngrp * per_block; q += ENUM_THREADS) {
                const int g = q / per_block, r = q - g * per_block;
                meta[g * C.spg + n_ground + b * per_block + r] = 0;
            }
        }
    }
    __syncthreads();
    const uint64_t dmask = (uint64_t)s_dmask[0] | ((uint64_t)s_dmask[1] << 32);     // rows with new pixels
    const bool env_full = n >= P.max_blocks;     // no placement possible: nothing is valid
    const int per_group = n_ground + s_nfree * n_offsets;
    const int total = s_ngrp * per_group;
    const int count = min(total, amax);
    if (tid == 0) {
        n_cand[e] = count;
        // generate_actions (actions.py:7-52) is unbounded: a list cut to the caller's capacity is reported
        // (bw_candidate_overflow), never dropped silently
        if (total > amax) atomicMax(P.cand_need, total);
    }
    // candidate a of the list: the action and the slot of its placement
    auto make_action = [&](int a, int &slot) {
        bw_action act;
        act.target_block = -1; act.target_face = 0; act.frozen = 0; act.reserved0 = 0; act.offset_y = 0.0;
        const int g = a / per_group, w = a - g * per_group;
        act.shape = s_grp_s[g];
        act.face = s_grp_f[g];
        if (w < n_ground) {
            act.offset_x = ground[w];
            slot = g * C.spg + w;
        } else {
            const int k = (w - n_ground) / n_offsets, oi = (w - n_ground) - k * n_offsets;
            act.target_block = s_free_b[k];
            act.target_face = s_free_f[k];
            act.offset_x = offsets[oi];
            slot = g * C.spg + n_ground + (act.target_block * NF + act.target_face) * n_offsets + oi;
        }
        return act;
    };
    if (env_full) {
        for (int a = tid; a < count; a += ENUM_THREADS) {
            int slot;
            cand[(size_t)e * amax + a] = make_action(a, slot);
            valid[(size_t)e * amax + a] = 0;
            if (slot_out != nullptr) slot_out[(size_t)e * amax + a] = -1;
        }
        if (COPY) {
            uint64_t *dst = action_bits + (size_t)e * amax * IMG;
            for (int q = tid; q < count * IMG; q += ENUM_THREADS) dst[q] = 0;
        }
        if (n_valid != nullptr && tid == 0) n_valid[e] = 0;
        return 0;
    }
    const double eps = 1e-6;
    const double xl = dsub(P.xlim0, eps), xh = dadd(P.xlim1, eps), zl = dsub(P.ylim0, eps), zh = dadd(P.ylim1, eps);

    // only the first `count` entries of the caller's buffers are meaningful (n_cand); the rest is left alone
    int myval = 0;
    for (int base = 0; base < count; base += CCH) {
        const int nch = min(CCH, count - base);
        // ---- A0
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            const int t = tid + j * ENUM_THREADS;
            bool is_miss = false, has_work = false;
            if (t < nch) {
                const int a = base + t;
                int slot;
                cand[(size_t)e * amax + a] = make_action(a, slot);
                const uint32_t m = meta[slot];
                uint32_t w = 0;
                int wlo = 0, wn = 0;
                if (m & SLOT_VALID) {
                    const int rows = slot_rows(m), ilo = slot_ilo(m);
                    const bool incr = !COPY && !fresh && ((m >> SLOT_STAMP_SHIFT) & SLOT_STAMP_MASK) == prev;
                    w = m & SLOT_GEOM;
                    if (incr) w |= (m & SLOT_OVL) | W_INCR;
                    if (COPY) {
                        wlo = ilo; wn = rows;
                    } else if (!(w & (SLOT_BAD | SLOT_OVL)) && rows > 0) {
                        if (incr) {
                            const uint64_t win = (rows >= 64) ? ~0ull : (((1ull << rows) - 1ull) << ilo);
                            const uint64_t sub = dmask & win;
                            if (sub) {
                                wlo = __ffsll((long long)sub) - 1;
                                wn = (63 - __clzll((long long)sub)) - wlo + 1;
                            }
                        } else {
                            wlo = ilo; wn = rows;
                        }
                    }
                    has_work = wn > 0;
                } else {
                    is_miss = true;
                }
                c_w[t] = w;
                c_slot[t] = slot;
                c_wlo[t] = (uint8_t)wlo;
                c_wn[t] = (uint8_t)wn;
            }
            const int im = list_append(is_miss, &s_nmiss, lane);
            if (is_miss) s_miss[im] = (uint16_t)t;
            const int iw = list_append(has_work, &s_nwork, lane);
            if (has_work) s_work[iw] = (uint16_t)t;
        }
        if (COPY) {   // rows outside the windows are zero
            uint64_t *dst = action_bits + ((size_t)e * amax + base) * IMG;
            for (int q = tid; q < nch * IMG; q += ENUM_THREADS) dst[q] = 0;
        }
        __syncthreads();
        const int nmiss = s_nmiss, nwork = s_nwork;
        // ---- Bh: 8 lanes per hit, two rows of a lane in flight
        for (int q = tid; q < nwork * 8; q += ENUM_THREADS) {
            const int t = s_work[q >> 3], r = q & 7;
            const int wlo = c_wlo[t], wend = wlo + c_wn[t];
            const uint64_t *tst = (c_w[t] & W_INCR) ? s_delta : s_full;
            const uint64_t *src = store + (size_t)c_slot[t] * IMG;
            uint64_t *dst = COPY ? action_bits + ((size_t)e * amax + base + t) * IMG : nullptr;
            bool ovl = false;
            for (int row = wlo + r; row < wend; row += 16) {
                const int row2 = row + 8;
                const bool two = row2 < wend;
                const uint64_t b0 = __ldcs(src + row);
                const uint64_t b1 = two ? __ldcs(src + row2) : 0ull;
                if (b0 & tst[row]) ovl = true;
                if (two && (b1 & tst[row2])) ovl = true;
                if (COPY) {
                    if (b0) dst[row] = b0;
                    if (b1) dst[row2] = b1;
                }
            }
            if (ovl) atomicOr(&c_w[t], SLOT_OVL);
        }
        // ---- misses, MISS_CAP per round
        int r0 = 0;
        do {
            const int nm = max(0, min(MISS_CAP, nmiss - r0));
            // ---- A1: eight lanes per miss -- lane j takes vertex j (bounds test, bounding box) and face j (posed
            // half-plane); the placement itself is computed by all eight
            for (int m0 = 0; m0 < nm; m0 += ENUM_THREADS / 8) {
                const int m = m0 + (tid >> 3), j = tid & 7;
                const bool on = m < nm;
                const int t = on ? s_miss[r0 + m] : 0;
                int slot = -1, err = 1;
                bw_action act;
                Pose ps;
                ps.x = ps.z = ps.s = 0.0; ps.c = 1.0;
                if (on) {
                    act = make_action(base + t, slot);
                    err = place_block(P, s_pose, s_shape, n, act, ps);
                }
                const bool placed = on && err == 0;
                const ShapeDev &sh = P.shapes[placed ? act.shape : 0];
                bool outside = false;
                double xmin = 1e300, xmax = -1e300, zmin = 1e300, zmax = -1e300;
                if (placed && j < sh.n_verts) {
                    double vx, vz;
                    rot(ps.c, ps.s, sh.vert_x[j], sh.vert_z[j], vx, vz);
                    vx = dadd(vx, ps.x);
                    vz = dadd(vz, ps.z);
                    // collision_on_action: any vertex outside the window (gym_env.py:304-323)
                    outside = vx < xl || vx > xh || vz < zl || vz > zh || vz < -eps;
                    xmin = xmax = vx;
                    zmin = zmax = vz;
                }
                const unsigned grp = (__ballot_sync(0xffffffffu, outside) >> (lane & 24)) & 0xffu;
#pragma unroll
                for (int o = 1; o < 8; o <<= 1) {
                    xmin = fmin(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
                    xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
                    zmin = fmin(zmin, __shfl_xor_sync(0xffffffffu, zmin, o));
                    zmax = fmax(zmax, __shfl_xor_sync(0xffffffffu, zmax, o));
                }
                if (placed && j < sh.n_faces) {          // pose_shape, face j
                    double fnx, fnz, ax, az;
                    rot(ps.c, ps.s, sh.face_nx[j], sh.face_nz[j], fnx, fnz);
                    rot(ps.c, ps.s, sh.face_cx[j], sh.face_cz[j], ax, az);
                    t_nx[m][j] = fnx;
                    t_nz[m][j] = fnz;
                    t_cx[m][j] = dadd(ax, ps.x);
                    t_cz[m][j] = dadd(az, ps.z);
                    t_inx[m][j] = (fnx != 0.0) ? 1.0 / fnx : 0.0;
                }
                if (on && j == 0) {
                    const bool bad = !placed || grp != 0;
                    int rows = 0, ilo = 0;
                    if (placed) {                        // conservative pixel window (pose_shape)
                        const int j_lo = max((int)floor((xmin - P.xlim0) * P.inv_step_x) - 1, 0);
                        const int j_hi = min((int)ceil((xmax - P.xlim0) * P.inv_step_x) + 1, IMG - 1);
                        const int i_lo = max((int)floor((P.ylim1 - zmax) * P.inv_step_y) - 1, 0);
                        const int i_hi = min((int)ceil((P.ylim1 - zmin) * P.inv_step_y) + 1, IMG - 1);
                        t_nf[m] = (int8_t)sh.n_faces;
                        t_jlo[m] = (int8_t)j_lo; t_jhi[m] = (int8_t)j_hi;
                        ilo = i_lo;
                        if (j_hi >= j_lo && i_hi >= i_lo) rows = i_hi - i_lo + 1;
                    } else {
                        c_slot[t] = -1;                  // a placement that failed is not kept
                    }
                    t_ilo[m] = (int8_t)ilo;
                    t_rows[m] = (int8_t)rows;
                    c_w[t] = (bad ? SLOT_BAD : 0u) | ((uint32_t)ilo << 7) | (uint32_t)rows;
                }
            }
            __syncthreads();
            if (tid == 0) { s_nmiss = 0; s_nwork = 0; }      // every thread has its copy; next use after the barrier in front of C
            // ---- Bm: sixteen lanes per miss, lane r takes the rows r, r + 16, ... of its window
            for (int q = tid; q < nm * 16; q += ENUM_THREADS) {
                const int mi = q >> 4;
                const int ilo = t_ilo[mi], iend = ilo + t_rows[mi];
                const int t = s_miss[r0 + mi];
                const int slot = c_slot[t];
                bool ovl = false;
                for (int row = ilo + (q & 15); row < iend; row += 16) {
                    const uint64_t bits = raster_row_posed_mixed(P, t_nf[mi], t_nx[mi], t_nz[mi], t_cx[mi], t_cz[mi],
                                                                 t_inx[mi], t_jlo[mi], t_jhi[mi], row);
                    if (slot >= 0) C.bits[((size_t)e * C.slots + slot) * IMG + row] = bits;
                    if (bits & s_full[row]) ovl = true;
                    if (COPY && bits) action_bits[((size_t)e * amax + base + t) * IMG + row] = bits;
                }
                if (ovl) atomicOr(&c_w[t], SLOT_OVL);
            }
            r0 += MISS_CAP;
            if (r0 < nmiss) __syncthreads();                 // the tables are written again
        } while (r0 < nmiss);
        __syncthreads();
        // ---- C
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            const int t = tid + j * ENUM_THREADS;
            if (t < nch) {
                const uint32_t w = c_w[t];
                const int slot = c_slot[t];
                const bool keep = !(w & (SLOT_BAD | SLOT_OVL));
                valid[(size_t)e * amax + base + t] = keep ? 1 : 0;
                myval += keep ? 1 : 0;
                if (slot_out != nullptr) slot_out[(size_t)e * amax + base + t] = slot;
                if (slot >= 0) meta[slot] = SLOT_VALID | (w & (SLOT_GEOM | SLOT_OVL)) | (cur << SLOT_STAMP_SHIFT);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) myval += __shfl_xor_sync(0xffffffffu, myval, o);
    if (lane == 0 && myval) atomicAdd(&s_nval, myval);
    __syncthreads();
    const int nval = s_nval;
    if (n_valid != nullptr && tid == 0) n_valid[e] = nval;
    return nval;
}

// One CTA per environment.  In the fused rollout (fin_stuck given) the CTA also closes the iteration: the transition
// that led here learns how many candidates its next state has, done |= "no candidate left" (successor_dqn.py:409-411),
// and an environment left without a candidate is restarted and enumerated once more by the same CTA.
template <bool COPY, bool FIN>
__global__ void __launch_bounds__(ENUM_THREADS, FIN ? 7 : 8)      // 7 CTAs per SM hold 1,024 environments in one wave
enumerate_store_kernel(Params PG, const double *__restrict__ ground, int n_ground, const double *__restrict__ offsets,
                       int n_offsets, int amax, bw_action *__restrict__ cand, uint8_t *__restrict__ valid,
                       int32_t *__restrict__ n_cand, uint64_t *__restrict__ action_bits, int32_t *__restrict__ slot_out,
                       CandCache C, const uint8_t *__restrict__ mask, int32_t *__restrict__ n_valid, RollFuse F) {
    const int e = blockIdx.x;
    if (mask != nullptr && mask[e] == 0) return;
    const int tid = threadIdx.x;
    bw_transition *const fin_slots = F.slots;
    uint8_t *const fin_stuck = F.R.stuck;
    // block library and pixel nodes in shared memory; the helpers of bw_common.cuh read them through P
    __shared__ __align__(16) unsigned char s_lib[BW_MAX_SHAPES * sizeof(ShapeDev)];
    __shared__ double s_grid[2 * IMG];
    __shared__ int s_stuck;
    Params P = PG;
    {
        const uint64_t *src = reinterpret_cast<const uint64_t *>(PG.shapes);
        uint64_t *dst = reinterpret_cast<uint64_t *>(s_lib);
        const int words = PG.n_shapes * (int)(sizeof(ShapeDev) / 8);
        for (int q = tid; q < words; q += ENUM_THREADS) dst[q] = src[q];
        if (tid < IMG) { s_grid[tid] = PG.xs[tid]; s_grid[IMG + tid] = PG.ys[tid]; }
        P.shapes = reinterpret_cast<const ShapeDev *>(s_lib);
        P.xs = s_grid;
        P.ys = s_grid + IMG;
    }
    if (!FIN) {
        enumerate_store_env<COPY>(P, e, ground, n_ground, offsets, n_offsets, amax, cand, valid, n_cand, action_bits,
                                  slot_out, C, n_valid);
        return;
    }
    if (F.out != nullptr) {   // rollout_record_kernel for this environment
        rollout_record_env(P, F.R, F.out, fin_slots, e, tid);
        const bool finished = P.done[e] != 0;
        __syncthreads();                  // every thread has read the flag the restart clears
        if (finished) restart_env(P, e, tid);
        __syncthreads();
    }
    int nv = 0;
    for (int pass = 0; pass < 2; pass++) {
        nv = enumerate_store_env<COPY>(P, e, ground, n_ground, offsets, n_offsets, amax, cand, valid, n_cand,
                                       action_bits, slot_out, C, n_valid);
        if (tid == 0) {       // rollout_finalize_kernel for this environment
            bool stuck = false;
            if (fin_slots != nullptr) {
                bw_transition &T = fin_slots[e];
                if (T.valid && !T.done) {
                    T.n_next_candidates = nv;
                    if (nv == 0) { T.done = 1; stuck = true; }
                } else if (!T.valid && nv == 0) {
                    stuck = true;
                }
            } else if (nv == 0) {
                stuck = true;
            }
            // a fresh environment always has its ground candidates; one without any is not restarted again
            if (stuck && P.n_blocks[e] == 0) stuck = false;
            fin_stuck[e] = stuck ? 1 : 0;
            s_stuck = stuck ? 1 : 0;
        }
        __syncthreads();
        if (!s_stuck) break;
        restart_env(P, e, tid);
        __syncthreads();
    }
    // rollout_pick_kernel (random policy) for the next iteration: the list is this CTA's own, nv its valid count
    if (F.next_slots != nullptr) rollout_pick_env(P, F.R, C, nullptr, 1, F.seed, F.next_step, F.next_slots, e, tid, nv);
}

void launch_enumerate(const Params &P, const double *d_ground, int n_ground, const double *d_offsets, int n_offsets,
                      int amax, bw_action *d_cand, uint8_t *d_valid, int32_t *d_n_cand, uint64_t *d_action_bits,
                      int32_t *d_slot, const CandCache &cache, cudaStream_t stream, const uint8_t *d_mask,
                      int32_t *d_n_valid, const RollFuse *fuse) {
    const RollFuse none;
    if (cache.meta != nullptr && cache.slots > 0) {
        if (d_action_bits != nullptr)
            enumerate_store_kernel<true, false><<<P.E, ENUM_THREADS, 0, stream>>>(
                P, d_ground, n_ground, d_offsets, n_offsets, amax, d_cand, d_valid, d_n_cand, d_action_bits, d_slot, cache,
                d_mask, d_n_valid, none);
        else if (fuse == nullptr)
            enumerate_store_kernel<false, false><<<P.E, ENUM_THREADS, 0, stream>>>(
                P, d_ground, n_ground, d_offsets, n_offsets, amax, d_cand, d_valid, d_n_cand, nullptr, d_slot, cache, d_mask,
                d_n_valid, none);
        else
            enumerate_store_kernel<false, true><<<P.E, ENUM_THREADS, 0, stream>>>(
                P, d_ground, n_ground, d_offsets, n_offsets, amax, d_cand, d_valid, d_n_cand, nullptr, d_slot, cache, d_mask,
                d_n_valid, *fuse);
    } else {
        enumerate_kernel<<<P.E, ENUM_THREADS, 0, stream>>>(P, d_ground, n_ground, d_offsets, n_offsets, amax, d_cand,
                                                           d_valid, d_n_cand, d_action_bits, d_mask, d_n_valid);
    }
}

// rasters of chosen candidates out of the store (or out of a dense [E,amax,64] copy): thread per (item, row)
__global__ void gather_bits_kernel(CandCache C, const int32_t *__restrict__ slot, const uint64_t *__restrict__ dense,
                                   int amax, int E, const int32_t *__restrict__ env, const int32_t *__restrict__ index,
                                   int64_t n, uint64_t *__restrict__ out) {
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n * IMG) return;
    const int64_t i = q / IMG;
    const int row = (int)(q - i * IMG);
    const int e = env ? env[i] : (int)i;
    const int a = index[i];
    uint64_t bits = 0;
    if (e >= 0 && e < E && a >= 0 && a < amax) {
        if (dense != nullptr) {
            bits = dense[((size_t)e * amax + a) * IMG + row];
        } else {
            const int s = slot[(size_t)e * amax + a];
            if (s >= 0 && s < C.slots) bits = cand_store_row(C, e, s, row);
        }
    }
    out[q] = bits;
}

void launch_gather_bits(const CandCache &cache, const int32_t *d_slot, const uint64_t *d_dense, int amax, int E,
                        const int32_t *d_env, const int32_t *d_index, int64_t n, uint64_t *d_out, cudaStream_t stream) {
    if (n <= 0) return;
    const int64_t threads = n * IMG;
    gather_bits_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(cache, d_slot, d_dense, amax, E, d_env, d_index,
                                                                              n, d_out);
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

// One warp per environment: a lane looks at 16 flags at a time (512 per trip of the warp).
__global__ void __launch_bounds__(128)
select_random_kernel(Params P, const bw_action *__restrict__ cand, const uint8_t *__restrict__ valid,
                     const int32_t *__restrict__ n_cand, int amax, uint64_t seed, bw_action *__restrict__ actions,
                     int32_t *__restrict__ index) {
    const int e = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (e >= P.E) return;
    const int cnt = n_cand[e];
    const uint8_t *row = valid + (size_t)e * amax;
    const bool wide = (amax & 15) == 0 && (reinterpret_cast<uintptr_t>(valid) & 15) == 0;
    int nvalid = 0;
    for (int base = lane * 16; base < cnt; base += 512) {
        uint32_t w[4];
        load_flags16(row, base, cnt, wide, w);
        nvalid += __popc(w[0]) + __popc(w[1]) + __popc(w[2]) + __popc(w[3]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
    int chosen = -1;
    if (nvalid > 0) {
        // the k-th valid candidate, k uniform (counter-based hash of seed, environment and its block count)
        const uint64_t r = mix64(seed ^ mix64((uint64_t)e * 0x632BE59BD9B4E019ull + (uint64_t)P.n_blocks[e]));
        chosen = kth_valid_candidate(row, cnt, wide, (int)(r % (uint64_t)nvalid), lane);
    }
    if (lane == 0) {
        bw_action act;
        act.target_block = -1; act.target_face = 0; act.shape = -1; act.face = 0;
        act.offset_x = 0.0; act.offset_y = 0.0; act.frozen = 0; act.reserved0 = 0;
        if (chosen >= 0) {
            act = cand[(size_t)e * amax + chosen];
        } else {
            // no candidate left: the episode ends here (rollout_episode, successor_dqn.py:409-411); the no-op
            // action leaves the flag alone and the next bw_reset_done starts the environment afresh
            P.done[e] = 1;
        }
        actions[e] = act;
        if (index != nullptr) index[e] = chosen;
    }
}

void launch_select_random(const Params &P, const bw_action *d_cand, const uint8_t *d_valid, const int32_t *d_n_cand,
                          int amax, uint64_t seed, bw_action *d_actions, int32_t *d_index, cudaStream_t stream) {
    select_random_kernel<<<(P.E + 3) / 4, 128, 0, stream>>>(P, d_cand, d_valid, d_n_cand, amax, seed, d_actions, d_index);
}

}  // namespace bw

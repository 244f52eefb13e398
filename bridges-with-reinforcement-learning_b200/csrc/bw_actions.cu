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

constexpr int ENUM_THREADS = 128;
constexpr int ENUM_CHUNK = 64;        // candidates posed per pass (shared-memory tables)

// Two phases per chunk of candidates.  A: one THREAD per candidate does everything that is uniform
// for the candidate (placement, bounds test, posed half-planes, pixel window) -- FP64 work that must
// not be replicated over the lanes of a warp.  B: one thread per (candidate, image row inside its
// window) evaluates the row's bit mask and the overlap with the block / obstacle rasters.
//
// CACHED: a candidate is (group, ground offset) or (group, target block, target face, offset) -- its pose, its
// bounds flag and its raster depend on nothing but the block library and the pose of the target block, so they
// survive from call to call (CandCache: one slot per possible candidate of an environment).  A block whose pose
// or shape differs from the copy taken when its slots were filled (a reset, a new block) invalidates its slots
// at the start of the call; ground slots live until the library or the offset tables change (host side).
// Only the overlap with the current block / obstacle rasters is recomputed for a cached candidate.
constexpr uint32_t SLOT_VALID = 0x80000000u, SLOT_BAD = 0x40000000u;

__device__ __forceinline__ bool same_bits(double a, double b) {
    return __double_as_longlong(a) == __double_as_longlong(b);
}

template <bool CACHED>
__global__ void __launch_bounds__(ENUM_THREADS, 8)
enumerate_kernel(Params PG, const double *__restrict__ ground, int n_ground, const double *__restrict__ offsets,
                 int n_offsets, int amax, bw_action *__restrict__ cand, uint8_t *__restrict__ valid,
                 int32_t *__restrict__ n_cand, uint64_t *__restrict__ action_bits, CandCache C,
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
    // A chunk is CHUNK consecutive candidates.  Only candidates that have to be posed and rasterised ("misses":
    // all of them without the cache) need a row of the posed-face tables; a chunk ends early when it would
    // hold more than MISS_CAP of them.
    constexpr int CHUNK = CACHED ? 2 * ENUM_CHUNK : ENUM_CHUNK;
    constexpr int MISS_CAP = ENUM_CHUNK;
    constexpr int CPL = CHUNK / 32;             // candidates per lane in the row-count scan
    __shared__ double c_nx[MISS_CAP][NF], c_nz[MISS_CAP][NF], c_cx[MISS_CAP][NF], c_cz[MISS_CAP][NF],
        c_inx[MISS_CAP][NF];
    __shared__ int8_t c_nf[MISS_CAP], c_jlo[MISS_CAP], c_jhi[MISS_CAP];
    __shared__ int8_t c_ilo[CHUNK], c_bad[CHUNK];
    __shared__ uint8_t c_mi[CHUNK];             // row of the posed-face tables (misses)
    __shared__ int c_rowstart[CHUNK + 1];
    __shared__ int c_overlap[CHUNK];
    __shared__ int c_slot[CHUNK];               // cache slot of the candidate (-1: none)
    __shared__ uint8_t c_cached[CHUNK];         // its slot was valid: raster rows are read, not computed
    __shared__ int s_wmiss[ENUM_THREADS / 32], s_ncut;
    constexpr int OWNER_CAP = 2048;             // (candidate, row) pairs of a chunk with a direct owner entry
    __shared__ uint8_t c_owner[OWNER_CAP];      // pair -> candidate of the chunk
    __shared__ unsigned s_inval;

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
    if (CACHED) {
        // blocks that are not the ones their slots were filled for: drop those slots, remember the new block
        if (tid == 0) s_inval = 0;
        __syncthreads();
        if (tid < n) {
            const Pose cp = C.pose[(size_t)e * NB + tid];
            const Pose p = s_pose[tid];
            const bool same = same_bits(cp.x, p.x) && same_bits(cp.z, p.z) && same_bits(cp.c, p.c) &&
                              same_bits(cp.s, p.s) && C.shape[(size_t)e * NB + tid] == s_shape[tid];
            if (!same) {
                C.pose[(size_t)e * NB + tid] = p;
                C.shape[(size_t)e * NB + tid] = s_shape[tid];
                atomicOr(&s_inval, 1u << tid);
            }
        }
        __syncthreads();
        unsigned inv = s_inval;
        const int per_block = NF * n_offsets;
        while (inv) {
            const int b = __ffs(inv) - 1;
            inv &= inv - 1;
            for (int q = tid; q < s_ngrp * per_block; q += ENUM_THREADS) {
                const int g = q / per_block, r = q - g * per_block;
                C.meta[(size_t)e * C.slots + g * C.spg + n_ground + b * per_block + r] = 0;
            }
        }
        __syncthreads();
    }
    const bool env_full = n >= P.max_blocks;     // no placement possible: nothing is cached, nothing is valid
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
    const int lane = tid & 31, warp = tid >> 5;
    int base = 0, nval = 0;
    while (base < count) {
        const int nch = min(CHUNK, count - base);
        // ---- phase A0: thread per candidate -- the action, its cache slot, hit or miss
        const bool active = tid < nch;
        bw_action act;
        act.target_block = -1; act.target_face = 0; act.frozen = 0; act.reserved0 = 0; act.offset_y = 0.0;
        act.shape = 0; act.face = 0; act.offset_x = 0.0;
        int slot = -1, rows = 0, ilo = 0;
        bool bad = false, hit = false;
        if (tid == 0) s_ncut = nch;
        if (active) {
            const int a = base + tid;
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
            cand[(size_t)e * amax + a] = act;
            if (!CACHED || env_full) slot = -1;
            if (CACHED && slot >= 0) {
                const uint32_t m = C.meta[(size_t)e * C.slots + slot];
                if (m & SLOT_VALID) {
                    hit = true;
                    bad = (m & SLOT_BAD) != 0;
                    rows = (int)(m & 0xffu);
                    ilo = (int)((m >> 8) & 0xffu);
                }
            }
        }
        // misses take the rows of the posed-face tables in candidate order; the chunk is cut in front of the
        // candidate that would need row MISS_CAP (it starts the next chunk)
        const bool miss = active && !hit;
        const unsigned mb = __ballot_sync(0xffffffffu, miss);
        if (lane == 0) s_wmiss[warp] = __popc(mb);
        __syncthreads();
        int before = __popc(mb & ((1u << lane) - 1u));
        for (int w = 0; w < warp; w++) before += s_wmiss[w];
        const bool included = active && (before + (miss ? 1 : 0) <= MISS_CAP);
        if (active && !included) atomicMin(&s_ncut, tid);
        // ---- phase A1: the misses are posed (FP64 work that is uniform for a candidate: one thread each)
        if (included) {
            if (miss) {
                const int mi = before;
                Pose ps;
                const int err = place_block(P, s_pose, s_shape, n, act, ps);
                bad = (err != 0);     // a full environment (err 2) offers no placement: listed but invalid
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
                        c_nx[mi][k] = o.nx[k]; c_nz[mi][k] = o.nz[k]; c_cx[mi][k] = o.cx[k]; c_cz[mi][k] = o.cz[k];
                        c_inx[mi][k] = o.inv_nx[k];
                    }
                    c_nf[mi] = (int8_t)o.n_faces;
                    c_jlo[mi] = (int8_t)o.j_lo; c_jhi[mi] = (int8_t)o.j_hi;
                    ilo = o.i_lo;
                    if (o.j_hi >= o.j_lo && o.i_hi >= o.i_lo) rows = o.i_hi - o.i_lo + 1;
                }
                if (CACHED && slot >= 0) {
                    if (err == 0)
                        C.meta[(size_t)e * C.slots + slot] = SLOT_VALID | (bad ? SLOT_BAD : 0u) | ((uint32_t)ilo << 8) | (uint32_t)rows;
                    else
                        slot = -1;
                }
                c_mi[tid] = (uint8_t)mi;
            }
            c_ilo[tid] = (int8_t)ilo;
            c_slot[tid] = slot;
            c_cached[tid] = hit ? 1 : 0;
            c_bad[tid] = bad ? 1 : 0;
            c_overlap[tid] = 0;
            c_rowstart[tid + 1] = rows;
        }
        __syncthreads();
        const int nchunk = s_ncut;                  // candidates base .. base + nchunk - 1 are finished in this trip
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
        // two pairs per trip: the loads of cached rows (global memory, possibly HBM) of both are in flight
        // before either is used
        for (int q0 = tid; q0 < npairs; q0 += 2 * ENUM_THREADS) {
            int tt[2], rr[2];
            uint64_t bb[2];
            bool live[2], hitv[2];
#pragma unroll
            for (int u = 0; u < 2; u++) {
                const int q = q0 + u * ENUM_THREADS;
                live[u] = q < npairs;
                hitv[u] = false;
                tt[u] = 0; rr[u] = 0; bb[u] = 0;
                if (live[u]) {
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
                    tt[u] = lo;
                    rr[u] = c_ilo[lo] + (q - c_rowstart[lo]);
                    hitv[u] = CACHED && c_cached[lo];
                    if (hitv[u]) bb[u] = __ldcs(&C.bits[((size_t)e * C.slots + c_slot[lo]) * IMG + rr[u]]);
                }
            }
#pragma unroll
            for (int u = 0; u < 2; u++) {
                if (!live[u]) continue;
                const int t = tt[u], row = rr[u];
                uint64_t bits = bb[u];
                if (!hitv[u]) {
                    const int mi = c_mi[t];
                    bits = raster_row_posed_mixed(P, c_nf[mi], c_nx[mi], c_nz[mi], c_cx[mi], c_cz[mi], c_inx[mi],
                                                  c_jlo[mi], c_jhi[mi], row);
                    if (CACHED && c_slot[t] >= 0) C.bits[((size_t)e * C.slots + c_slot[t]) * IMG + row] = bits;
                }
                if (bits & (s_block[row] | s_obst[row])) atomicOr(&c_overlap[t], 1);
                if (action_bits != nullptr && bits) action_bits[((size_t)e * amax + base + t) * IMG + row] = bits;
            }
        }
        __syncthreads();
        const bool keep = tid < nchunk && !c_bad[tid] && !c_overlap[tid];
        if (tid < nchunk) valid[(size_t)e * amax + base + tid] = keep ? 1 : 0;
        nval += __syncthreads_count(keep);
        base += nchunk;
    }
    if (n_valid != nullptr && tid == 0) n_valid[e] = nval;
}

void launch_enumerate(const Params &P, const double *d_ground, int n_ground, const double *d_offsets, int n_offsets,
                      int amax, bw_action *d_cand, uint8_t *d_valid, int32_t *d_n_cand, uint64_t *d_action_bits,
                      const CandCache &cache, cudaStream_t stream, const uint8_t *d_mask, int32_t *d_n_valid) {
    if (cache.meta != nullptr && cache.slots > 0)
        enumerate_kernel<true><<<P.E, ENUM_THREADS, 0, stream>>>(P, d_ground, n_ground, d_offsets, n_offsets, amax, d_cand,
                                                                 d_valid, d_n_cand, d_action_bits, cache, d_mask, d_n_valid);
    else
        enumerate_kernel<false><<<P.E, ENUM_THREADS, 0, stream>>>(P, d_ground, n_ground, d_offsets, n_offsets, amax, d_cand,
                                                                  d_valid, d_n_cand, d_action_bits, cache, d_mask, d_n_valid);
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
    const uint8_t *row = valid + (size_t)e * amax;
    // valid flags are 0/1 bytes: count / search them 16 at a time when the row is 16-byte aligned
    const bool wide = (amax & 15) == 0 && (reinterpret_cast<uintptr_t>(valid) & 15) == 0;
    int nvalid = 0;
    if (wide) {
        const uint4 *row4 = reinterpret_cast<const uint4 *>(row);
        for (int q = 0; q * 16 < cnt; q++) {
            uint4 v = row4[q];
            const int left = cnt - q * 16;               // flags past n_cand are stale: mask them off
            if (left < 16) {
                uint32_t w[4] = {v.x, v.y, v.z, v.w};
                for (int k = 0; k < 4; k++) {
                    const int keep = left - 4 * k;
                    if (keep <= 0) w[k] = 0;
                    else if (keep < 4) w[k] &= (1u << (8 * keep)) - 1;
                }
                v = make_uint4(w[0], w[1], w[2], w[3]);
            }
            nvalid += __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) +
                      __popc(v.w & 0x01010101u);
        }
    } else {
        for (int a = 0; a < cnt; a++) nvalid += row[a];
    }
    bw_action act;
    act.target_block = -1; act.target_face = 0; act.shape = -1; act.face = 0;
    act.offset_x = 0.0; act.offset_y = 0.0; act.frozen = 0; act.reserved0 = 0;
    int chosen = -1;
    if (nvalid > 0) {
        const uint64_t r = mix64(seed ^ mix64((uint64_t)e * 0x632BE59BD9B4E019ull + (uint64_t)P.n_blocks[e]));
        int k = (int)(r % (uint64_t)nvalid);
        int a = 0;
        if (wide) {                                       // skip whole 16-flag groups first
            const uint4 *row4 = reinterpret_cast<const uint4 *>(row);
            for (;; a += 16) {
                const uint4 v = row4[a >> 4];
                const int c16 = __popc(v.x & 0x01010101u) + __popc(v.y & 0x01010101u) + __popc(v.z & 0x01010101u) +
                                __popc(v.w & 0x01010101u);
                if (k < c16 || a + 16 >= cnt) break;
                k -= c16;
            }
        }
        for (; a < cnt; a++) {
            if (row[a]) {
                if (k == 0) { chosen = a; break; }
                k--;
            }
        }
        act = cand[(size_t)e * amax + chosen];
    } else {
        // no candidate left: the episode ends here (rollout_episode, successor_dqn.py:409-411); the no-op
        // action leaves the flag alone and the next bw_reset_done starts the environment afresh
        P.done[e] = 1;
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

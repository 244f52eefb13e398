// K4: reset-time rasters and the observation tensors.
//
//   reset_kernel    AssemblyGym.reset gym_env.py:255-289 + get_task_features
//                   successor_dqn.py:67-85 (obstacle raster, Gaussian-blurred target raster,
//                   convolve_with_gaussian robotoddler/utils/utils.py:93-114)
//   observe_kernel  get_state_features successor_dqn.py:47-64 / render_blocks_2d
//                   rendering.py:105-113: bit raster -> f32 [E,1,64,64], binary features [E,6]
//
// Rasters are kept bit-packed in HBM (512 B per image); the f32 expansion is a pure
// streaming write (16 KB per image) with 16-byte coalesced stores.
#include "bw_common.cuh"
#include "bw_kernels.cuh"

namespace bw {

constexpr int KSIZE = 101;   // successor_dqn.py:78-79
__constant__ float c_gauss[KSIZE];
__constant__ ShapeDev c_marker;   // cube06: obstacles and target markers (gym_env.py:277, successor_dqn.py:73)

void upload_obs_tables(const float *gauss, const ShapeDev *marker) {
    if (gauss) cudaMemcpyToSymbol(c_gauss, gauss, sizeof(float) * KSIZE);
    if (marker) cudaMemcpyToSymbol(c_marker, marker, sizeof(ShapeDev));
}

__global__ void __launch_bounds__(64)
reset_kernel(Params P, const bw_task *__restrict__ tasks, const uint8_t *__restrict__ mask, int only_done) {
    const int e = blockIdx.x;
    if (mask != nullptr && mask[e] == 0) return;
    const bool skip = only_done && P.done[e] == 0;
    __syncthreads();                      // every thread has read the flag the reset clears
    if (skip) return;
    const int tid = threadIdx.x;
    if (tasks == nullptr) {               // the task is kept: bw_reset_done
        restart_env(P, e, tid);
        return;
    }
    __shared__ TaskDev tk;
    __shared__ float s_tmp[IMG][IMG + 1];
    __shared__ uint64_t s_tbits[IMG];
    __shared__ Pose s_pose[NB];
    __shared__ uint8_t s_shape[NB];
    __shared__ int s_n;

    if (tid == 0) {
        {
            const bw_task &t = tasks[e];
            tk.n_obstacles = min(max(t.n_obstacles, 0), BW_MAX_OBSTACLES);
            tk.n_targets = min(max(t.n_targets, 0), BW_MAX_TARGETS);
            for (int i = 0; i < BW_MAX_OBSTACLES; i++) {
                tk.obstacle_xz[i][0] = t.obstacle_xz[i][0];
                tk.obstacle_xz[i][1] = t.obstacle_xz[i][1];
            }
            for (int i = 0; i < BW_MAX_TARGETS; i++) {
                tk.target_xz[i][0] = t.target_xz[i][0];
                tk.target_xz[i][1] = t.target_xz[i][1];
            }
            // a task this handle cannot hold (more pre-placed blocks than max_steps leaves room for, a shape
            // index outside the library) is refused as a whole: the environment starts empty and the call
            // that synchronises next reports it (bw_reset_host / bw_sync)
            int nb = max(t.n_blocks, 0);
            bool ok = nb <= P.max_blocks && t.n_obstacles >= 0 && t.n_obstacles <= BW_MAX_OBSTACLES &&
                      t.n_targets >= 0 && t.n_targets <= BW_MAX_TARGETS;
            for (int i = 0; ok && i < nb; i++) ok = t.blocks[i].shape >= 0 && t.blocks[i].shape < P.n_shapes;
            if (!ok) { nb = 0; atomicAdd(P.reset_err, 1); }
            uint32_t sm = 0;
            for (int i = 0; i < nb; i++) {
                Pose ps;
                ps.x = t.blocks[i].x; ps.z = t.blocks[i].z; ps.c = t.blocks[i].c; ps.s = t.blocks[i].s;
                s_pose[i] = ps;
                s_shape[i] = (uint8_t)t.blocks[i].shape;
                if (t.blocks[i].is_static) sm |= 1u << i;
                P.pose[(size_t)e * NB + i] = ps;
                P.shape_of[(size_t)e * NB + i] = (uint8_t)t.blocks[i].shape;
            }
            s_n = nb;
            P.static_mask[e] = sm;
        }
        for (int i = 0; i < BW_MAX_TARGETS; i++) { tk.remaining[i] = (int8_t)i; tk.reached[i] = -1; }
        tk.n_remaining = (int8_t)tk.n_targets;
        tk.n_reached = 0;
        P.task[e] = tk;
        P.n_blocks[e] = s_n;
        P.done[e] = 0;
        bw_step_out o;
        memset(&o, 0, sizeof(o));
        o.stable = 1;                     // empty assembly: stability.py:53-56
        o.stable_unfrozen = 1;
        for (int i = 0; i < BW_MAX_TARGETS; i++) o.distance_to_targets[i] = INFINITY;
        P.last_out[e] = o;
        P.su_valid[e] = (s_n == 0);       // pre-placed blocks: the evaluation that follows sets it
        P.warm_ok[2 * e] = 0;
        P.warm_ok[2 * e + 1] = 0;
        if (P.lp_meta != nullptr) {       // LpMeta (16 bytes) all zero: no basis
            reinterpret_cast<unsigned long long *>(P.lp_meta)[2 * e] = 0ull;
            reinterpret_cast<unsigned long long *>(P.lp_meta)[2 * e + 1] = 0ull;
        }
    }
    if (tid < NB) P.face_occ[(size_t)e * NB + tid] = 0;
    __syncthreads();

    // rasters: one thread per image row
    const int row = tid;
    uint64_t bbits = 0;
    for (int i = 0; i < s_n; i++) bbits |= raster_row(P, P.shapes[s_shape[i]], s_pose[i], row);
    P.block_bits[(size_t)e * IMG + row] = bbits;

    uint64_t obits = 0, tbits = 0;
    for (int i = 0; i < tk.n_obstacles; i++) {
        Pose ps;
        ps.x = tk.obstacle_xz[i][0]; ps.z = tk.obstacle_xz[i][1]; ps.c = 1.0; ps.s = 0.0;
        obits |= raster_row(P, c_marker, ps, row);
    }
    for (int i = 0; i < tk.n_targets; i++) {
        Pose ps;
        ps.x = tk.target_xz[i][0]; ps.z = tk.target_xz[i][1]; ps.c = 1.0; ps.s = 0.0;
        tbits |= raster_row(P, c_marker, ps, row);
    }
    P.obst_bits[(size_t)e * IMG + row] = obits;
    s_tbits[row] = tbits;
    // horizontal pass: tmp[row][j] = sum_v k[v - j + 50] * in[row][v]
    for (int j = 0; j < IMG; j++) {
        float acc = 0.0f;
        uint64_t bb = tbits;
        while (bb) {
            const int v = __ffsll((long long)bb) - 1;
            bb &= bb - 1;
            const int k = v - j + KSIZE / 2;
            if (k >= 0 && k < KSIZE) acc += c_gauss[k];
        }
        s_tmp[row][j] = acc;
    }
    __syncthreads();
    // vertical pass, thread = column
    const int col = tid;
    float *dst = P.reward_img + (size_t)e * IMG * IMG;
    for (int i = 0; i < IMG; i++) {
        float acc = 0.0f;
        for (int u = 0; u < IMG; u++) {
            if (s_tbits[u] == 0) continue;
            const int k = u - i + KSIZE / 2;
            if (k >= 0 && k < KSIZE) acc += c_gauss[k] * s_tmp[u][col];
        }
        dst[i * IMG + col] = acc;
    }
}

void launch_reset(const Params &P, const bw_task *d_tasks, const uint8_t *d_mask, int only_done, cudaStream_t stream) {
    reset_kernel<<<P.E, 64, 0, stream>>>(P, d_tasks, d_mask, only_done);
}

// render_blocks_2d for an explicit list of posed blocks (P.xs / P.ys give the pixel nodes)
__global__ void __launch_bounds__(64)
render_blocks_kernel(Params P, const ShapeDev *__restrict__ shapes, const bw_block *__restrict__ blocks, int n_blocks,
                     uint64_t *__restrict__ bits) {
    const int row = threadIdx.x;
    uint64_t acc = 0;
    for (int i = 0; i < n_blocks; i++) {
        Pose ps;
        ps.x = blocks[i].x; ps.z = blocks[i].z; ps.c = blocks[i].c; ps.s = blocks[i].s;
        acc |= raster_row(P, shapes[blocks[i].shape], ps, row);
    }
    bits[row] = acc;
}

// Shape.contains_2d (assembly_env.py:126-137) for arbitrary points: inside iff every half-plane value
// (p - c).n, evaluated with individually rounded operations like the rasters, is <= 0
__global__ void contains_points_kernel(ShapeDev sh, Pose ps, const double *__restrict__ pts, int64_t n,
                                       uint8_t *__restrict__ inside) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double px = pts[2 * i], pz = pts[2 * i + 1];
    bool in = true;
    for (int k = 0; k < sh.n_faces; k++) {
        double nx, nz, ax, az;
        rot(ps.c, ps.s, sh.face_nx[k], sh.face_nz[k], nx, nz);
        rot(ps.c, ps.s, sh.face_cx[k], sh.face_cz[k], ax, az);
        const double cx = dadd(ax, ps.x), cz = dadd(az, ps.z);
        in = in && (dadd(dmul(dsub(px, cx), nx), dmul(dsub(pz, cz), nz)) <= 0.0);
    }
    inside[i] = in ? 1 : 0;
}

void launch_contains_points(const ShapeDev &sh, const Pose &ps, const double *d_pts, int64_t n, uint8_t *d_inside,
                            cudaStream_t stream) {
    if (n <= 0) return;
    contains_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(sh, ps, d_pts, n, d_inside);
}

void launch_render_blocks(const Params &P, const ShapeDev *d_shapes, const bw_block *d_blocks, int n_blocks,
                          uint64_t *d_bits, cudaStream_t stream) {
    render_blocks_kernel<<<1, 64, 0, stream>>>(P, d_shapes, d_blocks, n_blocks, d_bits);
}

// ------------------------------------------------------------------ observation tensors
__device__ __forceinline__ float4 nibble_to_float4(uint64_t rowbits, int c4) {
    const unsigned nib = (unsigned)(rowbits >> (4 * c4)) & 0xfu;
    return make_float4((nib & 1u) ? 1.0f : 0.0f, (nib & 2u) ? 1.0f : 0.0f, (nib & 4u) ? 1.0f : 0.0f,
                       (nib & 8u) ? 1.0f : 0.0f);
}

// grid-stride over float4 elements of [n,64,64]; 16 float4 per image row
__global__ void __launch_bounds__(256)
expand_bits_kernel(const uint64_t *__restrict__ bits, int64_t n_vec4, float4 *__restrict__ img) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec4; i += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t rowbits = __ldg(bits + (i >> 4));
        __stcs(img + i, nibble_to_float4(rowbits, (int)(i & 15)));
    }
}

__global__ void __launch_bounds__(256)
copy_f4_kernel(const float4 *__restrict__ src, int64_t n_vec4, float4 *__restrict__ dst) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec4; i += (int64_t)gridDim.x * blockDim.x)
        __stcs(dst + i, __ldcs(src + i));
}

__global__ void binary_kernel(Params P, float *__restrict__ binary) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= P.E) return;
    const bw_step_out &o = P.last_out[e];
    float *b = binary + (size_t)e * 6;
    b[0] = o.stable; b[1] = o.collision; b[2] = o.collision_block;
    b[3] = o.collision_obstacle; b[4] = o.collision_floor; b[5] = o.collision_boundary;
}

static int grid_for(int64_t n_items, int threads) {
    int64_t blocks = (n_items + threads - 1) / threads;
    const int64_t cap = 148LL * 16;          // a few waves of the 148 SMs; grid-stride beyond that
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

void launch_expand_bits(const uint64_t *d_bits, int64_t n, float *d_img, cudaStream_t stream) {
    const int64_t n4 = n * (IMG * IMG / 4);
    expand_bits_kernel<<<grid_for(n4, 256), 256, 0, stream>>>(d_bits, n4, reinterpret_cast<float4 *>(d_img));
}

void launch_observe(const Params &P, float *d_block_img, float *d_binary, float *d_obstacle_img, float *d_reward_img,
                    cudaStream_t stream) {
    if (d_block_img) launch_expand_bits(P.block_bits, P.E, d_block_img, stream);
    if (d_obstacle_img) launch_expand_bits(P.obst_bits, P.E, d_obstacle_img, stream);
    if (d_reward_img) {
        const int64_t n4 = (int64_t)P.E * (IMG * IMG / 4);
        copy_f4_kernel<<<grid_for(n4, 256), 256, 0, stream>>>(reinterpret_cast<const float4 *>(P.reward_img), n4,
                                                              reinterpret_cast<float4 *>(d_reward_img));
    }
    if (d_binary) binary_kernel<<<(P.E + 127) / 128, 128, 0, stream>>>(P, d_binary);
}

// ------------------------------------------------------------------ FP64 FMA micro-benchmark
// Denominator of the solver's compute roofline (MEASURED_PEAKS.json has no FP64 entry):
// 8 independent FMA chains per thread, 1024 threads per SM-resident CTA pair.
__global__ void __launch_bounds__(512) fp64_fma_kernel(double *sink, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) sink[0] = s;
}

double measure_fp64_gflops(cudaStream_t stream) {
    double *sink = nullptr;
    if (cudaMalloc(&sink, 8) != cudaSuccess) return 0.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = 148 * 4, threads = 512, iters = 1 << 15;
    double best = 0.0;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0, stream);
        fp64_fma_kernel<<<blocks, threads, 0, stream>>>(sink, iters, 1.0000001, 1e-9);
        cudaEventRecord(e1, stream);
        cudaEventSynchronize(e1);
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double gflops = 2.0 * 8.0 * (double)iters * blocks * threads / (ms * 1e-3) * 1e-9;
        if (rep > 0 && gflops > best) best = gflops;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return best;
}

}  // namespace bw

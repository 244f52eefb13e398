// STUDY, NOT PART OF THE PRODUCT (not compiled into libbridges_b200.so): one warp per equilibrium system runs
// Lawson-Hanson NNLS on the friction-cone edge rays with an incremental orthogonal factorisation -- the CUDA
// form of oracle/nnls.py ("lead for the next round").  The program reads systems dumped by
// tools/dump_systems.py, solves each on the GPU, compares the residual with the file's (oracle) value and prints
// the cycles a lone warp needs per solve.  First (and so far only) run, profiles/r1_nnls_warp_study.txt: all 400
// fixture systems correct on a B200; untuned, about 17 k cycles per passive-set solve.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o nnls_warp nnls_warp.cu && ./nnls_warp systems.bin
//
// Lane = matrix row (m <= 32): row i of Q^T and of the triangle U live in shared memory with an odd stride, so a
// lane walks its own row or all lanes walk one column without bank conflicts.  Columns of R are sparse (a contact
// ray touches two blocks: <= 6 rows).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

constexpr int MR = 32;        // rows (3 per free block)
constexpr int MC = 128;       // columns (2 rays per contact point)
constexpr int NNZ = 6;
constexpr int LD = MR + 1;
constexpr unsigned FULL = 0xffffffffu;

struct Sys {                  // one system in device memory
    int m, n;
    double b[MR];
    unsigned char nnz[MC];
    unsigned char idx[MC][NNZ];
    double val[MC][NNZ];
};

__device__ __forceinline__ double wsum(double v) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ double wmin(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ double wmax(double v) {
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

__global__ void __launch_bounds__(32) nnls_kernel(const Sys *__restrict__ sys, int nsys, double *__restrict__ r_out,
                                                  int *__restrict__ it_out, long long *__restrict__ cyc_out) {
    const int e = blockIdx.x;
    if (e >= nsys) return;
    const int lane = threadIdx.x;
    __shared__ double Qt[MR * LD], U[MR * LD], qb[MR], res[MR], su[MR], invd[MR], xo[MR];
    __shared__ unsigned char order[MR], inP[MC], barred[MC];
    const Sys &S = sys[e];
    const int m = S.m, n = S.n;
    for (int q = lane; q < MR * LD; q += 32) { Qt[q] = 0.0; U[q] = 0.0; }
    for (int q = lane; q < MC; q += 32) { inP[q] = 0; barred[q] = 0; }
    __syncwarp();
    Qt[lane * LD + lane] = 1.0;
    qb[lane] = (lane < m) ? S.b[lane] : 0.0;
    xo[lane] = 0.0;
    __syncwarp();
    double scale = 0.0;
    for (int j = lane; j < n; j += 32)
        for (int t = 0; t < S.nnz[j]; t++) scale = fmax(scale, fabs(S.val[j][t]));
    scale = fmax(1.0, wmax(scale));
    const double tol = 1e-11 * scale;
    int p = 0, iters = 0;
    const long long t0 = clock64();
    const int max_iter = 6 * n + 50;
    while (iters < max_iter && p < m) {
        // residual of the passive-set solution: Q [0; (Q^T b)_tail]
        {
            double r = 0.0;
            if (lane < m)
                for (int i = p; i < m; i++) r += Qt[i * LD + lane] * qb[i];
            res[lane] = r;
        }
        __syncwarp();
        // gradient over the columns outside the passive set, warp arg-max
        double best = -INFINITY;
        int bj = -1;
        for (int j = lane; j < n; j += 32) {
            if (inP[j] || barred[j]) continue;
            double w = 0.0;
            for (int t = 0; t < S.nnz[j]; t++) w += S.val[j][t] * res[S.idx[j][t]];
            if (w > best) { best = w; bj = j; }
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(FULL, best, o);
            const int oj = __shfl_xor_sync(FULL, bj, o);
            if (ob > best || (ob == best && oj >= 0 && (bj < 0 || oj < bj))) { best = ob; bj = oj; }
        }
        if (!(best > tol) || bj < 0) break;
        // ---- the column enters: v = Q^T a, one Householder reflection on rows p..m-1
        double v = 0.0;
        if (lane < m)
            for (int t = 0; t < S.nnz[bj]; t++) v += S.val[bj][t] * Qt[lane * LD + S.idx[bj][t]];
        const double vv = wsum(v * v);
        const double tail2 = wsum((lane >= p && lane < m) ? v * v : 0.0);
        const double norm = sqrt(tail2);
        if (norm <= 1e-12 * fmax(1.0, sqrt(vv))) {           // in the span of the passive columns: not now
            if (lane == 0) barred[bj] = 1;
            __syncwarp();
            continue;
        }
        const double vp = __shfl_sync(FULL, v, p);
        const double alpha = (vp >= 0.0) ? -norm : norm;
        double u = (lane >= p && lane < m) ? v - (lane == p ? alpha : 0.0) : 0.0;
        const double un2 = wsum(u * u);
        if (un2 > 0.0) {
            u *= rsqrt(un2);
            su[lane] = u;
            __syncwarp();
            if (lane < m) {                                   // lane = column k of Q^T
                double t = 0.0;
                for (int i = p; i < m; i++) t += su[i] * Qt[i * LD + lane];
                t *= 2.0;
                for (int i = p; i < m; i++) Qt[i * LD + lane] -= su[i] * t;
            }
            const double dot = 2.0 * wsum((lane < m) ? u * qb[lane] : 0.0);
            if (lane >= p && lane < m) qb[lane] -= u * dot;
        }
        if (lane < p) U[lane * LD + p] = v;
        if (lane == p) { U[p * LD + p] = alpha; invd[p] = 1.0 / alpha; order[p] = (unsigned char)bj; inP[bj] = 1; xo[p] = 0.0; }
        for (int j = lane; j < n; j += 32) barred[j] = 0;
        p++;
        __syncwarp();
        // ---- least squares on the passive set; walk back while a coefficient would turn non-positive
        while (true) {
            iters++;
            double acc = (lane < p) ? qb[lane] : 0.0, s = 0.0;
            for (int k = p - 1; k >= 0; k--) {                // column-oriented back substitution
                const double sk = __shfl_sync(FULL, acc, k) * invd[k];
                if (lane == k) s = sk;
                if (lane < k) acc -= U[lane * LD + k] * sk;
            }
            const double smin = wmin((lane < p) ? s : INFINITY);
            if (p == 0 || smin > 0.0) {
                if (lane < p) xo[lane] = s;
                __syncwarp();
                break;
            }
            double x = (lane < p) ? xo[lane] : 0.0;
            const bool neg = (lane < p) && s <= 0.0;
            const double ratio = neg ? x / (x - s) : INFINITY;
            const double a = wmin(ratio);
            if (lane < p) x += a * (s - x);
            const double xmax = wmax((lane < p) ? fabs(x) : 0.0);
            unsigned drop = __ballot_sync(FULL, neg && x <= 1e-15 * fmax(1.0, xmax));
            if (!drop) {                                      // rounding left every candidate above the threshold:
                const unsigned cand = __ballot_sync(FULL, neg && ratio == a);   // the one that set the step leaves
                drop = cand & (0u - cand);
            }
            if (lane < p) xo[lane] = x;
            __syncwarp();
            while (drop) {
                const int k = 31 - __clz(drop);               // highest factor position first
                drop &= ~(1u << k);
                // delete column k of U (and its entries of order / xo): shift left
                const unsigned char col = order[k];
                if (lane < m) {                               // lane = row: left to right, each entry read before it is overwritten
#pragma unroll 1
                    for (int c = k; c < p - 1; c++) U[lane * LD + c] = U[lane * LD + c + 1];
                    U[lane * LD + p - 1] = 0.0;
                }
                const double xn = (lane + 1 < p) ? xo[lane + 1] : 0.0;
                const unsigned char on = (lane + 1 < p) ? order[lane + 1] : 0;
                __syncwarp();
                if (lane >= k && lane < p) { xo[lane] = xn; order[lane] = on; }
                if (lane == 0) inP[col] = 0;
                __syncwarp();
                // restore the triangle: Givens rotations on rows (i, i + 1)
                for (int i = k; i < p - 1; i++) {
                    const double ga = U[i * LD + i], gb = U[(i + 1) * LD + i];
                    double c = 1.0, sn = 0.0;
                    if (gb != 0.0) { const double rr = sqrt(ga * ga + gb * gb); c = ga / rr; sn = gb / rr; }
                    __syncwarp();
                    if (lane >= i && lane < p - 1) {
                        const double a0 = U[i * LD + lane], a1 = U[(i + 1) * LD + lane];
                        U[i * LD + lane] = c * a0 + sn * a1;
                        U[(i + 1) * LD + lane] = (lane == i) ? 0.0 : c * a1 - sn * a0;
                    }
                    if (lane < m) {
                        const double q0 = Qt[i * LD + lane], q1 = Qt[(i + 1) * LD + lane];
                        Qt[i * LD + lane] = c * q0 + sn * q1;
                        Qt[(i + 1) * LD + lane] = c * q1 - sn * q0;
                    }
                    if (lane == 0) {
                        const double b0 = qb[i], b1 = qb[i + 1];
                        qb[i] = c * b0 + sn * b1;
                        qb[i + 1] = c * b1 - sn * b0;
                    }
                    __syncwarp();
                    if (lane == 0) invd[i] = 1.0 / U[i * LD + i];
                }
                p--;
                __syncwarp();
            }
        }
    }
    const long long t1 = clock64();
    const double r2 = wsum((lane >= p && lane < m) ? qb[lane] * qb[lane] : 0.0);
    if (lane == 0) { r_out[e] = sqrt(r2); it_out[e] = iters; cyc_out[e] = t1 - t0; }
}

// file: int32 count; per system int32 m, n; double b[m] (normalised); double R[n][m] (column after column);
// double expected residual
int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s systems.bin\n", argv[0]); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror("open"); return 2; }
    int count = 0;
    if (fread(&count, 4, 1, f) != 1) return 2;
    std::vector<Sys> host;
    std::vector<double> expect;
    int skipped = 0;
    for (int s = 0; s < count; s++) {
        int m, n;
        if (fread(&m, 4, 1, f) != 1 || fread(&n, 4, 1, f) != 1) return 2;
        std::vector<double> b(m), R((size_t)m * n);
        double want;
        if (fread(b.data(), 8, m, f) != (size_t)m || fread(R.data(), 8, (size_t)m * n, f) != (size_t)m * n ||
            fread(&want, 8, 1, f) != 1) return 2;
        if (m > MR || n > MC || m == 0 || n == 0) { skipped++; continue; }
        Sys S;
        memset(&S, 0, sizeof(S));
        S.m = m; S.n = n;
        for (int i = 0; i < m; i++) S.b[i] = b[i];
        bool ok = true;
        for (int j = 0; j < n && ok; j++) {
            int k = 0;
            for (int i = 0; i < m; i++) {
                const double v = R[(size_t)j * m + i];
                if (v != 0.0) {
                    if (k == NNZ) { ok = false; break; }
                    S.idx[j][k] = (unsigned char)i; S.val[j][k] = v; k++;
                }
            }
            S.nnz[j] = (unsigned char)k;
        }
        if (!ok) { skipped++; continue; }
        host.push_back(S);
        expect.push_back(want);
    }
    fclose(f);
    const int N = (int)host.size();
    printf("%d systems (%d skipped: larger than %d x %d or a column with more than %d entries)\n", N, skipped, MR, MC, NNZ);
    if (N == 0) return 0;
    Sys *d_sys; double *d_r; int *d_it; long long *d_cyc;
    cudaMalloc(&d_sys, sizeof(Sys) * N); cudaMalloc(&d_r, 8 * N); cudaMalloc(&d_it, 4 * N); cudaMalloc(&d_cyc, 8 * N);
    cudaMemcpy(d_sys, host.data(), sizeof(Sys) * N, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    nnls_kernel<<<N, 32>>>(d_sys, N, d_r, d_it, d_cyc);            // warm-up
    cudaEventRecord(e0);
    nnls_kernel<<<N, 32>>>(d_sys, N, d_r, d_it, d_cyc);
    cudaEventRecord(e1);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { fprintf(stderr, "CUDA: %s\n", cudaGetErrorString(err)); return 1; }
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<double> r(N); std::vector<int> it(N); std::vector<long long> cyc(N);
    cudaMemcpy(r.data(), d_r, 8 * N, cudaMemcpyDeviceToHost);
    cudaMemcpy(it.data(), d_it, 4 * N, cudaMemcpyDeviceToHost);
    cudaMemcpy(cyc.data(), d_cyc, 8 * N, cudaMemcpyDeviceToHost);
    int bad = 0, flips = 0;
    double worst = 0.0, its = 0.0;
    for (int s = 0; s < N; s++) {
        const double d = fabs(r[s] - expect[s]);
        worst = std::max(worst, d);
        if (d > 1e-9 + 1e-8 * expect[s]) { if (bad < 10) printf("  system %d: residual %.3e, expected %.3e\n", s, r[s], expect[s]); bad++; }
        if ((r[s] <= 1e-6) != (expect[s] <= 1e-6)) flips++;
        its += it[s];
    }
    std::vector<long long> sorted(cyc);
    std::sort(sorted.begin(), sorted.end());
    double mean = 0; for (auto c : cyc) mean += (double)c; mean /= N;
    printf("residual mismatches %d (worst |diff| %.2e), verdict flips %d, passive-set solves mean %.1f\n", bad, worst, flips, its / N);
    printf("cycles per solve: mean %.0f  p50 %lld  p99 %lld  max %lld;  launch of %d warps: %.3f ms\n", mean, sorted[N / 2],
           sorted[(size_t)(0.99 * (N - 1))], sorted[N - 1], N, ms);
    return bad ? 1 : 0;
}

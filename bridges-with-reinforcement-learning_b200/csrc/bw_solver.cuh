// K3: the rigid-block equilibrium solver, one warp per problem, everything in shared memory.
//
// Replaces `rbe_solve(assembly, mu, density, penalty=False)` as called by `is_stable_rbe`
// (assembly_gym/assembly_gym/utils/stability.py:49-71).  Problem (2-D reduction):
//     find f in K = prod {(fn, ft): |ft| <= mu fn}  with  A f = b,
// three rows per free block.  Computed quantity: r = min_{f in K} ||A f - b|| (b normalised) and,
// when r = 0, the minimum-norm f.  Method: proximal-point iteration on the
// dual  d(y) = b.y - 1/2 ||P_K(A^T y)||^2  with rho = 1e4, 1e8, 1e8, ...; every proximal sub-problem by
// a semismooth Newton method:
//     gradient   b - A P_K(A^T y) - (y - y_k)/rho
//     Hessian    A J A^T + I/rho      (J = generalised Jacobian of the cone projection)
//     direction  packed left-looking Cholesky, lane = row; the right-hand side rides along as one
//                more row so that the forward substitution costs nothing
//     step       bracketing search on the piecewise-linear derivative phi'(t)
// Loops are deliberately kept rolled: the whole solver must stay resident in the instruction
// cache (fully unrolled variants were 3x slower, profiles/README.md).
#pragma once
#include "bw_common.cuh"

namespace bw {

constexpr unsigned FULL = 0xffffffffu;
constexpr int NSCHED = 6;
__constant__ double c_rho[NSCHED] = {1e4, 1e8, 1e8, 1e8, 1e8, 1e8};
constexpr int MAX_NEWTON = 60;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}
__device__ __forceinline__ void warp_sum2(double &a, double &b) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(FULL, a, o);
        b += __shfl_xor_sync(FULL, b, o);
    }
}
__device__ __forceinline__ void warp_sum3(double &a, double &b, double &c) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(FULL, a, o);
        b += __shfl_xor_sync(FULL, b, o);
        c += __shfl_xor_sync(FULL, c, o);
    }
}
__device__ __forceinline__ void warp_sum4(double &a, double &b, double &c, double &d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(FULL, a, o);
        b += __shfl_xor_sync(FULL, b, o);
        c += __shfl_xor_sync(FULL, c, o);
        d += __shfl_xor_sync(FULL, d, o);
    }
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ int tri(int i) { return (i * (i + 1)) >> 1; }

// 1/sqrt(a) for a normal positive a: hardware approximation + one Newton step (~1e-13 relative),
// ample for the Newton direction (the gradient, which fixes the solution, does not use it)
__device__ __forceinline__ double fast_rsqrt(double a) {
    double x;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    const double e = fma(-a * x, x, 1.0);
    return fma(0.5 * x, e, x);
}

// 1/a for a normal a: hardware approximation + one Newton step (line-search abscissae only)
__device__ __forceinline__ double fast_rcp(double a) {
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(a));
    return fma(fma(-a, x, 1.0), x, x);
}

// projection of (gn, gt) onto the friction cone |ft| <= mu fn.
// typ: 0 = polar cone (f = 0), 1 = interior, 2 = ray ft = +mu fn, 3 = ray ft = -mu fn
__device__ __forceinline__ void project_cone(double gn, double gt, double mu, double inv_den, double &fn, double &ft,
                                             int &typ) {
    // written with selects: the contact points of a warp sit on different faces, and three serialised
    // code paths cost more than the few redundant operations
    const double agt = fabs(gt);
    const bool inter = agt <= mu * gn;
    const bool polar = mu * agt <= -gn;
    const double k = (gn + mu * agt) * inv_den;
    const double kt = (gt > 0) ? mu * k : -(mu * k);
    fn = inter ? gn : (polar ? 0.0 : k);
    ft = inter ? gt : (polar ? 0.0 : kt);
    typ = inter ? 1 : (polar ? 0 : (gt > 0 ? 2 : 3));
}

// TWO = false: at most 31 rows (10 free blocks + the right-hand side), one row per lane;
// TWO = true: up to 49 rows, lanes also own row lane + 32.
template <bool TWO>
struct Solver {
    // shared by the two problems of an environment, read-only here
    const double *G;         // [nc][12]: per contact point the columns of A for body a (normal 0..2,
                             //           tangent 3..5) and body b (6..8, 9..11); rows = Fx, Fz, torque/L0
    const uint8_t *c_a, *c_b;   // bodies of a contact point (0 = floor)
    const uint8_t *adj_ptr;     // [NBODY+1] CSR over bodies
    const uint8_t *adj;         // contact | side << 7
    // this problem
    double *y, *yk, *d, *b, *g, *h, *f, *invd;
    double *L;               // packed lower triangle, rows 0..m; row m carries the right-hand side
    unsigned *adjm;          // [NBODY] scratch of the mechanism screen
    uint8_t *typ;
    int8_t *rowbase;         // body -> first row or -1 (support)
    uint8_t *freebody;       // free block index -> body
    uint8_t *firstcol;       // free block index -> first column of its rows' envelope (set by solve())
    int m, nfree, nc, nitf, lane;
    double mu, inv_den;
    double r_exit;           // a residual <= r_exit ends the solve as feasible (1e-9 when the forces are
                             // wanted, else the verdict threshold: the verdict cannot change below it)
    const volatile int *sibling;   // verdict of the other solve of this environment (-1 = still running)
    int implied_by;          // sibling verdict that decides this solve too (-1: never)
    bool exit_anytime;       // false when the min-norm forces are wanted (needs the converged dual iterate)
    double flops;            // work estimate, see DESIGN.md section 6
#ifdef BW_PROFILE
    long long acc_t[6];      // grad, assemble, factor+solve, A^T d + dots, line search + update, residual
    long long t_screen;      // mechanism screen
    long long acc_f[3];      // factor_and_solve: dot loops, pivot blocks + stores, back substitution
#define BW_T0(name) const long long name = clock64()
#define BW_ACC(i, t0) acc_t[i] += clock64() - (t0)
#else
#define BW_T0(name)
#define BW_ACC(i, t0)
#endif

    // out = A^T v (lanes over contact points)
    __device__ __forceinline__ void at_times(const double *v, double *out) const {
#pragma unroll 1
        for (int c = lane; c < nc; c += 32) {
            const double *Gc = G + c * 12;
            const int ra = rowbase[c_a[c]], rb = rowbase[c_b[c]];
            double gn = 0.0, gt = 0.0;
            if (ra >= 0) {
                const double v0 = v[ra], v1 = v[ra + 1], v2 = v[ra + 2];
                gn = Gc[0] * v0 + Gc[1] * v1 + Gc[2] * v2;
                gt = Gc[3] * v0 + Gc[4] * v1 + Gc[5] * v2;
            }
            if (rb >= 0) {
                const double v0 = v[rb], v1 = v[rb + 1], v2 = v[rb + 2];
                gn += Gc[6] * v0 + Gc[7] * v1 + Gc[8] * v2;
                gt += Gc[9] * v0 + Gc[10] * v1 + Gc[11] * v2;
            }
            out[2 * c] = gn;
            out[2 * c + 1] = gt;
        }
    }

    // f = P_K(g), typ; returns whether any contact point changed its cone face since the last call
    __device__ __forceinline__ bool project_all() {
        bool changed = false;
#pragma unroll 1
        for (int c = lane; c < nc; c += 32) {
            double fn, ft;
            int t;
            project_cone(g[2 * c], g[2 * c + 1], mu, inv_den, fn, ft, t);
            f[2 * c] = fn;
            f[2 * c + 1] = ft;
            changed |= (t != (int)typ[c]);
            typ[c] = (uint8_t)t;
        }
        return __any_sync(FULL, changed);
    }

    // (A f)_i
    __device__ __forceinline__ double a_times_f_row(int i) const {
        const int I = i / 3, k = i - 3 * I;
        const int body = freebody[I];
        double acc = 0.0;
#pragma unroll 2
        for (int q = adj_ptr[body]; q < adj_ptr[body + 1]; q++) {
            const int e = adj[q];
            const int c = e & 0x7f;
            const double *Gc = G + c * 12 + (e >> 7) * 6;
            acc += Gc[k] * f[2 * c] + Gc[3 + k] * f[2 * c + 1];
        }
        return acc;
    }

    // generalised Jacobian of the cone projection of a contact point from its face type, without branches:
    // J = alpha I + beta (1, s)(1, s)^T  with alpha = 1 in the interior, beta = 1/(1+mu^2), s = +-mu on the
    // two rays, J = 0 in the polar cone.  (jnn, jnt, jtt) = (J00, J01, J11).
    __device__ __forceinline__ void jac(int tp, double mu2, double &jnn, double &jnt, double &jtt) const {
        const double alpha = (tp == 1) ? 1.0 : 0.0;
        const double beta = (tp >= 2) ? inv_den : 0.0;
        const double sb = (tp == 3) ? -beta : beta;
        jnn = alpha + beta;
        jnt = mu * sb;
        jtt = fma(mu2, beta, alpha);
    }

    // One pass over the rows (lane = row, three lanes per free block) walks the contact list of the row's
    // block once for two results: the gradient of the proximal sub-problem  b - A f - (y - y_k)/rho  with the
    // equilibrium residual b - A f, and row r of the block's diagonal 3x3 tile of H = A J A^T + I/rho.
    // Every contact point runs the same instruction sequence whatever its cone face (the three faces of a
    // warp's contact points used to serialise three code paths).  Writes d = rhs row = gradient, zero-fills
    // rows 0..m-1 of L first; returns the partial sums of |grad|^2 and |b - A f|^2.
    __device__ __forceinline__ void rows_pass(double inv_rho, double *rhs, double &gn2, double &rr2) {
        const int nz = tri(m);
#pragma unroll 1
        for (int q = lane; q < nz; q += 32) L[q] = 0.0;
        __syncwarp();
        const double mu2 = mu * mu;
        gn2 = 0.0; rr2 = 0.0;
#pragma unroll 1
        for (int q = lane; q < m; q += 32) {
            const int I = q / 3, r = q - 3 * I;
            const int body = freebody[I];
            double e0 = 0.0, e1 = 0.0, e2 = 0.0;         // entries (r, 0..r) of the tile
            double acc = 0.0;                            // (A f)_q
            const int a1 = adj_ptr[body + 1];
#pragma unroll 2
            for (int a = adj_ptr[body]; a < a1; a++) {
                const int e = adj[a];
                const int c = e & 0x7f;
                const double *Gi = G + c * 12 + (e >> 7) * 6;
                const double n0 = Gi[0], n1 = Gi[1], n2 = Gi[2], t0 = Gi[3], t1 = Gi[4], t2 = Gi[5];
                const double nr = Gi[r], tr = Gi[3 + r];
                acc += nr * f[2 * c] + tr * f[2 * c + 1];
                double jnn, jnt, jtt;
                jac(typ[c], mu2, jnn, jnt, jtt);
                const double ur = fma(nr, jnn, tr * jnt), vr = fma(nr, jnt, tr * jtt);   // row r of G_i^T J
                e0 = fma(ur, n0, fma(vr, t0, e0));
                e1 = fma(ur, n1, fma(vr, t1, e1));
                e2 = fma(ur, n2, fma(vr, t2, e2));
            }
            const double px = (y[q] - yk[q]) * inv_rho;
            const double rs = b[q] - acc;
            const double gr = rs - px;
            d[q] = gr;                                   // kept for the dot products after the solve
            rhs[q] = gr;
            gn2 += gr * gr;
            rr2 += rs * rs;
            double *p = L + tri(q) + 3 * I;
            p[0] = e0 + (r == 0 ? inv_rho : 0.0);
            if (r >= 1) p[1] = e1 + (r == 1 ? inv_rho : 0.0);
            if (r == 2) p[2] = e2 + inv_rho;
        }
        warp_sum2(gn2, rr2);
    }

    // Off-diagonal 3x3 tiles of H: a tile belongs to exactly one interface (two contact points), so lanes
    // over interfaces write it without accumulation conflicts.
    __device__ __forceinline__ void offdiag_tiles() {
        const double mu2 = mu * mu;
#pragma unroll 1
        for (int k = lane; k < nitf; k += 32) {
            const int c0 = 2 * k;
            const int ra = rowbase[c_a[c0]], rb = rowbase[c_b[c0]];
            if (ra < 0 || rb < 0) continue;
            double t00 = 0, t01 = 0, t02 = 0, t10 = 0, t11 = 0, t12 = 0, t20 = 0, t21 = 0, t22 = 0;
#pragma unroll
            for (int q = 0; q < 2; q++) {
                const int c = c0 + q;
                const double *Ga = G + c * 12, *Gb = Ga + 6;
                double jnn, jnt, jtt;
                jac(typ[c], mu2, jnn, jnt, jtt);
                // W = J [Ga_n; Ga_t]  (2 x 3), tile += Gb_n^T W_n + Gb_t^T W_t
                const double wn0 = fma(jnn, Ga[0], jnt * Ga[3]), wn1 = fma(jnn, Ga[1], jnt * Ga[4]),
                             wn2 = fma(jnn, Ga[2], jnt * Ga[5]);
                const double wt0 = fma(jnt, Ga[0], jtt * Ga[3]), wt1 = fma(jnt, Ga[1], jtt * Ga[4]),
                             wt2 = fma(jnt, Ga[2], jtt * Ga[5]);
                const double u0 = Gb[0], u1 = Gb[1], u2 = Gb[2], v0 = Gb[3], v1 = Gb[4], v2 = Gb[5];
                t00 = fma(u0, wn0, fma(v0, wt0, t00)); t01 = fma(u0, wn1, fma(v0, wt1, t01)); t02 = fma(u0, wn2, fma(v0, wt2, t02));
                t10 = fma(u1, wn0, fma(v1, wt0, t10)); t11 = fma(u1, wn1, fma(v1, wt1, t11)); t12 = fma(u1, wn2, fma(v1, wt2, t12));
                t20 = fma(u2, wn0, fma(v2, wt0, t20)); t21 = fma(u2, wn1, fma(v2, wt1, t21)); t22 = fma(u2, wn2, fma(v2, wt2, t22));
            }
            double *p = L + tri(rb) + ra;          // rb > ra: rows of the later block
            p[0] = t00; p[1] = t01; p[2] = t02;
            p += rb + 1;
            p[0] = t10; p[1] = t11; p[2] = t12;
            p += rb + 2;
            p[0] = t20; p[1] = t21; p[2] = t22;
        }
        __syncwarp();
    }

    // Left-looking Cholesky of rows 0..m-1 with row m (the gradient) carried along, then the
    // back substitution: d = H^-1 grad.  The factorisation advances one 3x3 block column (= one free
    // block) at a time: every lane accumulates the three dot products of its row with the three rows
    // of the block (one own load feeds three FMAs), the 3x3 diagonal block is factorised redundantly
    // by all lanes from six shuffled values, and each lane finishes its three entries with a 3x3
    // triangular solve.  m is a multiple of 3, so the right-hand-side row m lies below every block.
    template <bool TW>
    __device__ void factor_and_solve(double inv_rho) {
        constexpr bool TWO = TW;                           // (shadows the class parameter inside this function)
        const int nrows = m + 1;
        const double sqrt_rho = fast_rsqrt(inv_rho);
        const int i0 = lane, i1 = lane + 32;
        double *row0 = L + tri(i0 < nrows ? i0 : 0);       // idle lanes read row 0 (results unused)
        double *row1 = L + tri((TWO && i1 < nrows) ? i1 : 0);
#pragma unroll 1
        for (int c0 = 0; c0 < m; c0 += 3) {
            const double *rA = L + tri(c0), *rB = L + tri(c0 + 1), *rC = L + tri(c0 + 2);
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, b0 = 0.0, b1 = 0.0, b2 = 0.0;
            BW_T0(t_f0);
            // c0 is a multiple of 3: three columns per trip, all twelve loads issued before the FMAs
            // (columns left of the envelope of the pivot rows hold exact zeros: skipped)
#pragma unroll 2
            for (int p = firstcol[c0 / 3]; p < c0; p += 3) {
                const double la0 = rA[p], lb0 = rB[p], lc0 = rC[p];
                const double la1 = rA[p + 1], lb1 = rB[p + 1], lc1 = rC[p + 1];
                const double la2 = rA[p + 2], lb2 = rB[p + 2], lc2 = rC[p + 2];
                const double o0 = row0[p], o1 = row0[p + 1], o2 = row0[p + 2];
                a0 += o0 * la0; a1 += o0 * lb0; a2 += o0 * lc0;
                a0 += o1 * la1; a1 += o1 * lb1; a2 += o1 * lc1;
                a0 += o2 * la2; a1 += o2 * lb2; a2 += o2 * lc2;
                if (TWO) {
                    const double q0 = row1[p], q1 = row1[p + 1], q2 = row1[p + 2];
                    b0 += q0 * la0; b1 += q0 * lb0; b2 += q0 * lc0;
                    b0 += q1 * la1; b1 += q1 * lb1; b2 += q1 * lc1;
                    b0 += q2 * la2; b1 += q2 * lb2; b2 += q2 * lc2;
                }
            }
            // t = H[i][c0..c0+2] - partial dots (entries right of the diagonal are never used)
            const double t0 = row0[c0] - a0, t1 = row0[c0 + 1] - a1, t2 = row0[c0 + 2] - a2;
#ifdef BW_PROFILE
            acc_f[0] += clock64() - t_f0;
            const long long t_f1 = clock64();
#endif
            double u0 = 0.0, u1 = 0.0, u2 = 0.0;
            if (TWO) { u0 = row1[c0] - b0; u1 = row1[c0 + 1] - b1; u2 = row1[c0 + 2] - b2; }
            // the diagonal block: rows c0, c0+1, c0+2
            const int ra = c0, rb = c0 + 1, rc = c0 + 2;
            const double d00 = __shfl_sync(FULL, (!TWO || ra < 32) ? t0 : u0, ra & 31);
            const double d10 = __shfl_sync(FULL, (!TWO || rb < 32) ? t0 : u0, rb & 31);
            const double d11 = __shfl_sync(FULL, (!TWO || rb < 32) ? t1 : u1, rb & 31);
            const double d20 = __shfl_sync(FULL, (!TWO || rc < 32) ? t0 : u0, rc & 31);
            const double d21 = __shfl_sync(FULL, (!TWO || rc < 32) ? t1 : u1, rc & 31);
            const double d22 = __shfl_sync(FULL, (!TWO || rc < 32) ? t2 : u2, rc & 31);
            // 3x3 Cholesky of the diagonal block through its leading minors: with s11, s22 the Schur
            // complements, M2 = d00 s11 and T = d00 M2 s22 need no earlier square root, so the three
            // rsqrt chains run side by side instead of one after the other
            //   1/sqrt(s11) = rsqrt(M2) sqrt(d00),   1/sqrt(s22) = rsqrt(T) sqrt(d00) sqrt(M2)
            // (H is positive definite; a pivot that rounding drives to <= 0 -- or NaN -- is replaced by 1/rho.
            // The tests run beside the rsqrt chains and only select the results.)
            const double i00r = fast_rsqrt(d00);
            const double M2 = fma(d11, d00, -(d10 * d10));
            const double c21 = fma(d21, d00, -(d20 * d10));            // d00 (d21 - l20 l10)
            const double T = fma(fma(d00, d22, -(d20 * d20)), M2, -(c21 * c21));   // d00 M2 s22
            const double rM2 = fast_rsqrt(M2), rT = fast_rsqrt(T);
            const bool ok0 = d00 > 1e-300, ok1 = ok0 && M2 > 1e-300 * d00, ok2 = ok1 && T > 1e-300 * (d00 * M2);
            const double sp0 = d00 * i00r, sM2 = M2 * rM2;             // sqrt(d00), sqrt(M2)
            const double i00 = ok0 ? i00r : sqrt_rho;
            const double i11 = ok1 ? rM2 * sp0 : sqrt_rho;
            const double i22 = ok2 ? rT * (sp0 * sM2) : sqrt_rho;
            const double l10 = d10 * i00, l20 = d20 * i00;
            const double l21 = (d21 - l20 * l10) * i11;
            // the strictly lower entries of the INVERSE of the diagonal block take the place of l10, l20, l21
            // (nothing reads those again but the back substitution, which then needs no triangular solve)
            const double m10 = -(l10 * i00) * i11;
            const double m21 = -(l21 * i11) * i22;
            const double m20 = -fma(l21, m10, l20 * i00) * i22;
            // x L_d^T = t for this lane's row(s)
            {
                const double x0 = t0 * i00;
                const double x1 = (t1 - x0 * l10) * i11;
                const double x2 = (t2 - x0 * l20 - x1 * l21) * i22;
                if (i0 < nrows) {
                    if (i0 > ra) row0[ra] = (i0 == rb) ? m10 : (i0 == rc ? m20 : x0);
                    if (i0 > rb) row0[rb] = (i0 == rc) ? m21 : x1;
                    if (i0 > rc) row0[rc] = x2;
                }
            }
            if (TWO && i1 < nrows) {
                const double x0 = u0 * i00;
                const double x1 = (u1 - x0 * l10) * i11;
                const double x2 = (u2 - x0 * l20 - x1 * l21) * i22;
                if (i1 > ra) row1[ra] = (i1 == rb) ? m10 : (i1 == rc ? m20 : x0);
                if (i1 > rb) row1[rb] = (i1 == rc) ? m21 : x1;
                if (i1 > rc) row1[rc] = x2;
            }
            if (lane == 0) { invd[ra] = i00; invd[rb] = i11; invd[rc] = i22; }
            __syncwarp();
#ifdef BW_PROFILE
            acc_f[1] += clock64() - t_f1;
#endif
        }
        BW_T0(t_f2);
        // row m now holds z = L^-1 grad; back substitution L^T d = z, one 3x3 block per trip: the block's
        // three unknowns are three short dot products with the stored inverse of the diagonal block, every
        // lane runs the same instructions (results picked by selects) and its loads do not wait for the chain
        const double *rowm = L + tri(m);
        double z0 = (i0 < m) ? rowm[i0] : 0.0;
        double z1 = (TWO && i1 < m) ? rowm[i1] : 0.0;
#pragma unroll 1
        for (int c0 = m - 3; c0 >= 0; c0 -= 3) {
            const int ra = c0, rb = c0 + 1, rc = c0 + 2;
            const double *rA = L + tri(ra), *rB = L + tri(rb), *rC = L + tri(rc);
            const int k0 = (i0 < ra) ? i0 : 0;
            const double la = rA[k0], lb = rB[k0], lc = rC[k0];
            const double m10 = rB[ra], m20 = rC[ra], m21 = rC[rb];
            const double i00 = invd[ra], i11 = invd[rb], i22 = invd[rc];
            const double za = __shfl_sync(FULL, (!TWO || ra < 32) ? z0 : z1, ra & 31);
            const double zb = __shfl_sync(FULL, (!TWO || rb < 32) ? z0 : z1, rb & 31);
            const double zc = __shfl_sync(FULL, (!TWO || rc < 32) ? z0 : z1, rc & 31);
            // (da, db, dc) = L_d^-T (za, zb, zc)
            const double dc = i22 * zc;
            const double db = fma(m21, zc, i11 * zb);
            const double da = fma(m20, zc, fma(m10, zb, i00 * za));
            {
                const double upd = fma(-la, da, fma(-lb, db, fma(-lc, dc, z0)));
                z0 = (i0 < ra) ? upd : (i0 == ra ? da : (i0 == rb ? db : (i0 == rc ? dc : z0)));
            }
            if (TWO) {
                const int k1 = (i1 < ra) ? i1 : 0;
                const double upd = fma(-rA[k1], da, fma(-rB[k1], db, fma(-rC[k1], dc, z1)));
                z1 = (i1 < ra) ? upd : (i1 == ra ? da : (i1 == rb ? db : (i1 == rc ? dc : z1)));
            }
        }
        if (i0 < m) d[i0] = z0;
        if (TWO && i1 < m) d[i1] = z1;
        __syncwarp();
#ifdef BW_PROFILE
        acc_f[2] += clock64() - t_f2;
#endif
    }

    // ||b - A P_K(A^T y)|| (b is normalised); leaves g = A^T y, f = P_K(g)
    __device__ double residual() {
        at_times(y, g);
        __syncwarp();
        project_all();
        __syncwarp();
        double acc = 0.0;
#pragma unroll 1
        for (int i = lane; i < m; i += 32) {
            const double r = b[i] - a_times_f_row(i);
            acc += r * r;
        }
        return sqrt(warp_sum(acc));
    }

    // sum over contact points of P_K(g + t h) . h
    __device__ __forceinline__ double fdoth(double t) {
        double fh = 0.0;
#pragma unroll 1
        for (int c = lane; c < nc; c += 32) {
            double fn, ft;
            int tp;
            const double hn = h[2 * c], ht = h[2 * c + 1];
            project_cone(g[2 * c] + t * hn, g[2 * c + 1] + t * ht, mu, inv_den, fn, ft, tp);
            fh += fn * hn + ft * ht;
        }
        flops += 20.0 * nc;
        return warp_sum(fh);
    }

    // Mechanism screen: a Farkas certificate of "no equilibrium" whose dual vector is one rigid virtual
    // motion of a sub-assembly S of the free blocks.  Inside S the relative motion at every contact is
    // zero, so only the contacts between S and the rest matter: with the wrenches (Fx, Fz, torque/L0 about
    // the origin) of the two friction-cone edge rays of every such contact point, sigma_k r_k (sigma = +1
    // when S holds body b of the contact, -1 when it holds body a), and the weight wrench b_S of S,
    //     n . (sigma_k r_k) >= 0 for all boundary rays  and  n . b_S < 0
    // proves that no non-negative combination of the rays balances b_S, i.e. A f = b has no solution in K.
    // Candidate sets: for every free block i the blocks resting on it, S_i = {i} + every later free block
    // in contact with a member (for a single tower: the top of the structure from i upwards), from the
    // last block down.  Candidate motions n per boundary contact point: the rotation about that point
    // (r+ x r-) and the translations perpendicular to its two edge rays, both signs.  Checked with margins
    // (EPS on the rays, DELTA on the weight) so that only clear certificates count; everything else goes
    // to solve().  On the bench rollouts this decides 96 % of the systems without equilibrium
    // (tools/solver_lab.py: an LP on the aggregated 3-row systems catches the same cases).
    // All sets are examined in one pass: lane = (set, boundary contact point) work item, at most a few
    // chunks of 32 items (a tower has two boundary contact points per set).
    // Scratch: g, h, f (rays), invd (bodies of a contact), adjm (contact adjacency masks of the bodies),
    // y, yk, d (work items).
    __device__ bool screen(const double *body, double invL0) {
        constexpr double EPS = 1e-10, DELTA = 1e-5;
        if (nc > 64) return false;
        uint8_t *cba = reinterpret_cast<uint8_t *>(invd), *cbb = cba + nc;
        if (lane < NBODY) adjm[lane] = 0u;
        __syncwarp();
#pragma unroll 1
        for (int c = lane; c < nc; c += 32) {
            const double *Gb = G + c * 12 + 6;                  // wrench on body b about its centroid
            const int A = c_a[c], B = c_b[c];
            const double cx = body[B * 8], cz = body[B * 8 + 1];
            const double nx = Gb[0], nz = Gb[1], nt = Gb[2] + (cx * Gb[1] - cz * Gb[0]) * invL0;
            const double tx = Gb[3], tz = Gb[4], tt = Gb[5] + (cx * Gb[4] - cz * Gb[3]) * invL0;
            const double px = nx + mu * tx, pz = nz + mu * tz, pt = nt + mu * tt;
            const double qx = nx - mu * tx, qz = nz - mu * tz, qt = nt - mu * tt;
            const double ip = fast_rsqrt(px * px + pz * pz + pt * pt), iq = fast_rsqrt(qx * qx + qz * qz + qt * qt);
            g[2 * c] = px * ip; g[2 * c + 1] = pz * ip; f[2 * c] = pt * ip;
            h[2 * c] = qx * iq; h[2 * c + 1] = qz * iq; f[2 * c + 1] = qt * iq;
            cba[c] = (uint8_t)A;
            cbb[c] = (uint8_t)B;
            if (rowbase[A] >= 0 && rowbase[B] >= 0) {
                atomicOr(&adjm[A], 1u << B);
                atomicOr(&adjm[B], 1u << A);
            }
        }
        __syncwarp();
        // lane = free block i: S_i as a mask over bodies and its weight wrench (0, ws, ts)
        int mybody = 0;
        unsigned myadj = 0, S = 0;
        double w1 = 0.0, t1 = 0.0;
        if (lane < nfree) {
            mybody = freebody[lane];
            myadj = adjm[mybody];
            S = 1u << mybody;
            w1 = b[3 * lane + 1];
            t1 = body[mybody * 8] * w1 * invL0;
        }
        double ws = w1, ts = t1;
#pragma unroll 1
        for (int j = 1; j < nfree; j++) {
            const int bj = __shfl_sync(FULL, mybody, j);
            const unsigned aj = __shfl_sync(FULL, myadj, j);
            const double wj = __shfl_sync(FULL, w1, j), tj = __shfl_sync(FULL, t1, j);
            if (j > lane && (aj & S)) { S |= 1u << bj; ws += wj; ts += tj; }
        }
        // boundary contacts of S_lane as bit masks over the contact points (bm) and, among them, those with S on
        // the a side (am: sigma = -1); all sets at once, lane = set
        unsigned bm0 = 0, bm1 = 0, am0 = 0, am1 = 0;
#pragma unroll 1
        for (int c = 0; c < nc; c++) {
            const unsigned ina = (S >> cba[c]) & 1u, inb = (S >> cbb[c]) & 1u;
            const unsigned bd = ina ^ inb, sa = ina & ~inb;
            if (c < 32) { bm0 |= bd << c; am0 |= sa << c; }
            else { bm1 |= bd << (c - 32); am1 |= sa << (c - 32); }
        }
        int nbd = __popc(bm0) + __popc(bm1);
        // nothing holds S
        if (__any_sync(FULL, lane < nfree && nbd == 0 && ws > 0.0)) return true;
        if (lane >= nfree || nbd > 32) nbd = 0;          // sets with more than 32 boundary contacts are not examined
        // one work item per (set, boundary contact point): the candidate motions of that contact point against
        // all boundary rays of its set.  Items are listed set by set (exclusive scan of the counts) in the
        // memory of y, yk, d, which solve() initialises later.
        int inc = nbd;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(FULL, inc, o);
            if (lane >= o) inc += v;
        }
        const int total = __shfl_sync(FULL, inc, 31);
        uint16_t *items = reinterpret_cast<uint16_t *>(y);
        const int cap = 12 * (int)(yk - y);               // 3 arrays of (yk - y) doubles, four items per double
        if (total == 0 || total > cap) return false;
        {
            int pos = inc - nbd;
            unsigned w0 = nbd ? bm0 : 0u, w1m = nbd ? bm1 : 0u;
#pragma unroll 1
            while (w0) {
                const int c = __ffs(w0) - 1;
                w0 &= w0 - 1;
                items[pos++] = (uint16_t)((lane << 8) | c);
            }
#pragma unroll 1
            while (w1m) {
                const int c = __ffs(w1m) - 1;
                w1m &= w1m - 1;
                items[pos++] = (uint16_t)((lane << 8) | (c + 32));
            }
        }
        __syncwarp();
        bool found = false;
#pragma unroll 1
        for (int base = 0; base < total; base += 32) {
            const int pidx = base + lane;
            const bool act = pidx < total;
            const int it = act ? items[pidx] : 0;
            const int set = it >> 8, q = it & 0xff;
            unsigned r0 = __shfl_sync(FULL, bm0, set), r1 = __shfl_sync(FULL, bm1, set);
            const unsigned s0 = __shfl_sync(FULL, am0, set), s1 = __shfl_sync(FULL, am1, set);
            const double bw = __shfl_sync(FULL, ws, set), bt = __shfl_sync(FULL, ts, set);
            if (act) {
                const double ax = g[2 * q], az = g[2 * q + 1], at = f[2 * q];
                const double ex = h[2 * q], ez = h[2 * q + 1], et = f[2 * q + 1];
                // rotation about the contact point; translations perpendicular to the two edge rays
                const double n0x = az * et - at * ez, n0z = at * ex - ax * et, n0t = ax * ez - az * ex;
                const double n1x = az, n1z = -ax, n2x = ez, n2z = -ex;
                // with l = |n|:  all sigma r . n >= -EPS l  and  n . b_S <= -DELTA l |b_S|   (or the mirror image).
                // Only the signs matter: per motion two flags "some ray below -EPS l" / "some ray above +EPS l"
                const double l0 = n0x * n0x + n0z * n0z + n0t * n0t, l1 = n1x * n1x + n1z * n1z, l2 = n2x * n2x + n2z * n2z;
                const double tol0 = EPS * (l0 * fast_rsqrt(fmax(l0, 1e-300))), tol1 = EPS * (l1 * fast_rsqrt(fmax(l1, 1e-300))),
                             tol2 = EPS * (l2 * fast_rsqrt(fmax(l2, 1e-300)));
                bool neg0 = false, pos0 = false, neg1 = false, pos1 = false, neg2 = false, pos2 = false;
#pragma unroll 1
                for (int half = 0; half < 2; half++) {
                    unsigned w = half ? r1 : r0;
                    const unsigned sm = half ? s1 : s0;
#pragma unroll 1
                    while (w) {
                        const int kb = __ffs(w) - 1;
                        w &= w - 1;
                        const int k = kb + 32 * half;
                        const double sg = ((sm >> kb) & 1u) ? -1.0 : 1.0;
                        const double rx = g[2 * k], rz = g[2 * k + 1], rt = f[2 * k];
                        const double sx = h[2 * k], sz = h[2 * k + 1], st = f[2 * k + 1];
                        const double d0 = sg * (n0x * rx + n0z * rz + n0t * rt), e0 = sg * (n0x * sx + n0z * sz + n0t * st);
                        const double d1 = sg * (n1x * rx + n1z * rz), e1 = sg * (n1x * sx + n1z * sz);
                        const double d2 = sg * (n2x * rx + n2z * rz), e2 = sg * (n2x * sx + n2z * sz);
                        neg0 |= (d0 < -tol0) | (e0 < -tol0); pos0 |= (d0 > tol0) | (e0 > tol0);
                        neg1 |= (d1 < -tol1) | (e1 < -tol1); pos1 |= (d1 > tol1) | (e1 > tol1);
                        neg2 |= (d2 < -tol2) | (e2 < -tol2); pos2 |= (d2 > tol2) | (e2 > tol2);
                    }
                }
                const double bn2 = bw * bw + bt * bt;
                const double w0 = n0z * bw + n0t * bt, w1n = n1z * bw, w2n = n2z * bw;   // n . b_S, b_S = (0, bw, bt)
                const double D2 = DELTA * DELTA * bn2;
                if (l0 > 1e-18) found |= (!neg0 && w0 < 0.0 && w0 * w0 >= D2 * l0) || (!pos0 && w0 > 0.0 && w0 * w0 >= D2 * l0);
                if (l1 > 1e-18) found |= (!neg1 && w1n < 0.0 && w1n * w1n >= D2 * l1) || (!pos1 && w1n > 0.0 && w1n * w1n >= D2 * l1);
                if (l2 > 1e-18) found |= (!neg2 && w2n < 0.0 && w2n * w2n >= D2 * l2) || (!pos2 && w2n > 0.0 && w2n * w2n >= D2 * l2);
            }
            if (__any_sync(FULL, found)) return true;
        }
        return false;
    }

    // returns status: 0 feasible (r <= r_exit), 1 stalled at r* > 0, 2 not converged,
    // 3 verdict implied by the sibling solve (released-block equilibrium => frozen-block equilibrium)
    // warm = true: y holds a starting point (the dual iterate of an earlier solve on (almost) the same rows,
    // see step_kernel); the proximal-point iteration converges from any start, a good one saves Newton steps.
    __device__ int solve(double &r_out, int &iters_out, bool warm = false) {
        if (!warm) {
#pragma unroll 1
            for (int i = lane; i < m; i += 32) y[i] = 0.0;
#pragma unroll 1
            for (int c = lane; c < 2 * nc; c += 32) g[c] = 0.0;  // g = A^T y is kept up to date
        } else {
            __syncwarp();
            at_times(y, g);
        }
#pragma unroll 1
        for (int c = lane; c < nc; c += 32) typ[c] = 255;
        // row envelope of H = A J A^T (and of its Cholesky factor, which only fills inside it): the rows of
        // free block I start at the rows of the earliest free block it touches
        if (lane < nfree) {
            const int body = freebody[lane];
            int first = 3 * lane;
#pragma unroll 1
            for (int q = adj_ptr[body]; q < adj_ptr[body + 1]; q++) {
                const int e = adj[q];
                const int other = (e >> 7) ? c_a[e & 0x7f] : c_b[e & 0x7f];
                const int ro = rowbase[other];
                if (ro >= 0 && ro < first) first = ro;
            }
            firstcol[lane] = (uint8_t)first;
        }
        __syncwarp();
        project_all();                         // f, cone faces of g = 0; kept up to date by every step below
        __syncwarp();
        flops = 0.0;
        double *rhs = L + tri(m);
        double rprev = -1.0, r = 1.0;
        int status = 2, iters = 0;
#pragma unroll 1
        for (int k = 0; k < NSCHED; k++) {
            const double inv_rho = 1.0 / c_rho[k];
#pragma unroll 1
            for (int i = lane; i < m; i += 32) yk[i] = y[i];
            __syncwarp();
            bool have_r = false;
            bool full_step = false;            // the previous Newton step of this stage was taken with t = 1
            bool changed = true;               // a contact changed its cone face in that step
#pragma unroll 1
            for (int it = 0; it < MAX_NEWTON; it++) {
                // one lane reads the flag and broadcasts it: the exit must be taken by the whole warp or not at
                // all (the code below is full of full-mask shuffles), and the sibling may publish between two
                // lanes' loads
                if (implied_by >= 0) {
                    int sv = (lane == 0) ? *sibling : 0;
                    sv = __shfl_sync(FULL, sv, 0);
                    if (sv == implied_by) { status = 3; break; }
                }
                BW_T0(t_a);
                double gn2, rr2;
                rows_pass(inv_rho, rhs, gn2, rr2);
                BW_ACC(0, t_a);
                // f is in K, so ||b - A f|| bounds r* from above at every iterate.  A full Newton step that
                // leaves every contact on its cone face has solved the (then quadratic) sub-problem exactly:
                // what is left of the gradient is rounding noise (it grows with |y| and can stay above the
                // absolute threshold for ever on systems without equilibrium)
                if (gn2 <= 1e-20 || (full_step && !changed) || (exit_anytime && rr2 <= r_exit * r_exit)) {
                    r = sqrt(rr2);
                    have_r = true;
                    break;
                }
                BW_T0(t_b);
                offdiag_tiles();
                double gd = 0.0;
                // keep the gradient in registers: d[] is overwritten by the solve
                const double gr0 = (lane < m) ? d[lane] : 0.0;
                const double gr1 = (TWO && lane + 32 < m) ? d[lane + 32] : 0.0;
                __syncwarp();
                BW_ACC(1, t_b);
                BW_T0(t_c);
                // systems of up to 31 rows (10 free blocks) take the one-row-per-lane form even in the 16-block
                // instantiation: the second row of a lane would be all predicated-off instructions
                if (TWO && m + 1 > 32) factor_and_solve<true>(inv_rho);
                else factor_and_solve<false>(inv_rho);
                BW_ACC(2, t_c);
                BW_T0(t_d);
                // phi'(t) = grad.d + f.h - P_K(g + t h).h - t d.d / rho   (piecewise linear, decreasing).
                // One pass over the contact points gives h = A^T d, f.h and the first evaluation P_K(g + h).h
                double dd = 0.0, fh0 = 0.0, fh1 = 0.0;
                {
                    const double d0 = (lane < m) ? d[lane] : 0.0;
                    gd = gr0 * d0;
                    dd = d0 * d0;
                    if (TWO) {
                        const double d1 = (lane + 32 < m) ? d[lane + 32] : 0.0;
                        gd += gr1 * d1;
                        dd += d1 * d1;
                    }
#pragma unroll 1
                    for (int c = lane; c < nc; c += 32) {
                        const double *Gc = G + c * 12;
                        const int ra = rowbase[c_a[c]], rb = rowbase[c_b[c]];
                        double hn = 0.0, ht = 0.0;
                        if (ra >= 0) {
                            const double v0 = d[ra], v1 = d[ra + 1], v2 = d[ra + 2];
                            hn = Gc[0] * v0 + Gc[1] * v1 + Gc[2] * v2;
                            ht = Gc[3] * v0 + Gc[4] * v1 + Gc[5] * v2;
                        }
                        if (rb >= 0) {
                            const double v0 = d[rb], v1 = d[rb + 1], v2 = d[rb + 2];
                            hn += Gc[6] * v0 + Gc[7] * v1 + Gc[8] * v2;
                            ht += Gc[9] * v0 + Gc[10] * v1 + Gc[11] * v2;
                        }
                        h[2 * c] = hn;
                        h[2 * c + 1] = ht;
                        fh0 += f[2 * c] * hn + f[2 * c + 1] * ht;
                        double fn, ft;
                        int tp;
                        project_cone(g[2 * c] + hn, g[2 * c + 1] + ht, mu, inv_den, fn, ft, tp);
                        fh1 += fn * hn + ft * ht;
                    }
                    flops += 20.0 * nc;
                }
                warp_sum4(gd, dd, fh0, fh1);
                BW_ACC(3, t_d);
                BW_T0(t_e);
                const double phi0 = gd;
                if (!(phi0 > 1e-30)) break;
                const double base = phi0 + fh0;
                double t = 1.0;
                double p = base - fh1 - dd * inv_rho;
                // the full step is kept whenever it satisfies the search's own acceptance test
                // |phi'(1)| <= 0.1 phi'(0): on a piece without a change of cone face phi'(1) is rounding noise
                // of either sign (it grows with |y|), and a search started by that noise ends at t = 0.9 --
                // the gradient then shrinks by only 10x per Newton step instead of vanishing
                if (p < -0.1 * phi0) {
                    // bracket the root with a safeguarded regula falsi until |phi'| <= 0.1 phi'(0)
                    double lo = 0.0, plo = phi0, hi = 1.0, phi = p;
#pragma unroll 1
                    for (int ls = 0; ls < 20; ls++) {
                        const double w = hi - lo;
                        t = lo + w * plo * fast_rcp(plo - phi);      // plo > 0 > phi; the safeguard below bounds t anyway
                        t = fmin(fmax(t, lo + 0.1 * w), hi - 0.1 * w);
                        p = base - fdoth(t) - t * dd * inv_rho;
                        if (fabs(p) <= 0.1 * phi0) break;
                        if (p > 0.0) { lo = t; plo = p; } else { hi = t; phi = p; }
                    }
                    if (p < 0.0 && fabs(p) > 0.1 * phi0 && lo > 0.0) t = lo;
                }
                double yy = 0.0;
#pragma unroll 1
                for (int i = lane; i < m; i += 32) {
                    const double yn = y[i] + t * d[i];
                    y[i] = yn;
                    yy = fmax(yy, fabs(yn));
                }
                // g = A^T (y + t d) and its projection f, cone faces for the next gradient pass
                {
                    bool ch = false;
#pragma unroll 1
                    for (int c = lane; c < nc; c += 32) {
                        const double gn = g[2 * c] + t * h[2 * c], gt = g[2 * c + 1] + t * h[2 * c + 1];
                        g[2 * c] = gn;
                        g[2 * c + 1] = gt;
                        double fn, ft;
                        int tp;
                        project_cone(gn, gt, mu, inv_den, fn, ft, tp);
                        f[2 * c] = fn;
                        f[2 * c + 1] = ft;
                        ch |= (tp != (int)typ[c]);
                        typ[c] = (uint8_t)tp;
                    }
                    changed = __any_sync(FULL, ch);
                }
                __syncwarp();
                BW_ACC(4, t_e);
                iters++;
                full_step = (t == 1.0);
                flops += (double)m * m * m / 3.0 + 2.0 * m * m + 156.0 * nc + 12.0 * m;
                // no representable progress any more (|t d| below the rounding of y)
                const double ymax = warp_max(yy);
                if (t * t * dd <= 1e-30 * fmax(1.0, ymax * ymax)) break;
            }
            if (status == 3) break;
            if (!have_r) {                         // left the Newton loop without a fresh gradient pass
                BW_T0(t_f);
                r = residual();
                BW_ACC(5, t_f);
            }
            if (r <= r_exit) { status = 0; break; }
            if (rprev >= 0.0 && fabs(r - rprev) <= 1e-3 * r) { status = 1; break; }
            // a feasible system loses two orders of magnitude per stage (rho x 100); a residual that
            // stays above 90% of its previous value and far above the verdict threshold has stalled at r*
            if (rprev >= 0.0 && r >= 0.9 * rprev && r > 1e-3) { status = 1; break; }
            rprev = r;
        }
        r_out = r;
        iters_out = iters;
        return status;
    }
};

}  // namespace bw

// K3b: the equilibrium VERDICT of a real step as a warm-started linear programme, one warp per problem.
//
// Same problem as bw_solver.cuh (`rbe_solve` behind `is_stable_rbe`, assembly_gym/utils/stability.py:49-71):
//     find f in K = prod {(fn, ft): |ft| <= mu fn}  with  A f = b.
// In the 2-D reduction K is polyhedral: f = lambda+ (1, +mu) + lambda- (1, -mu) per contact point, so the question
// is the feasibility of  R lambda = b, lambda >= 0  with the friction-cone edge rays  r(c, +-) = a_n(c) +- mu a_t(c)
// as columns.  Phase 1 of the revised simplex method answers it with a certificate either way:
//     feasible    a basic solution lambda >= 0; the forces f it stands for are checked against the contact data
//                 directly, ||b - A f|| <= stable_tol  (the same test that ends the Newton solve)
//     infeasible  the dual vector pi of the optimal basis: pi . r <= eps for EVERY ray and pi . b > delta, a
//                 Farkas certificate (checked against the contact data, not against the basis inverse)
// so a verdict never depends on the conditioning of the basis inverse that was used to find it: a run that cannot
// certify its answer returns LP_NONE and the problem goes to the Newton solver.
//
// What makes it fast is the history of a rollout.  A step adds one block: the released problem of step t is the
// released problem of step t-1 plus three rows (the new block's) and the new block's contact columns; the frozen
// problem of step t has exactly the rows of the released problem of step t-1 plus support columns.  The optimal
// basis of the released problem is therefore kept per environment in HBM (basis inverse, column identities) and
// both problems of the next step start from it: 2-3 pivots per solve on average, 30 at most on 90,000 harvested
// systems (tools/simplex_lab.py), where the semismooth Newton method walks 12-23 cone-face changes of ~15 k cycles
// on the hard ones.
//
// Column identity across steps: (body pair index, contact point 0/1, ray sign) -- interfaces of a body pair never
// change while both blocks stay where they are; artificial columns are +e_i (b >= 0: weights) and only ever leave.
#pragma once
#include "bw_solver.cuh"

namespace bw {

constexpr uint16_t LP_ART = 0x8000;
enum { LP_NONE = 0, LP_FEASIBLE = 1, LP_INFEASIBLE = 2 };

struct LpMeta {            // per environment, next to the stored basis
    uint32_t mask;         // free blocks (bit i = block i) whose rows the stored basis has, in block order
    uint16_t m;            // rows of the stored basis (3 per free block); 0 = no basis (always a valid start)
    uint16_t feasible;     // the stored basis is a feasible one (no artificial column above zero)
    double L0;             // torque-row scale (largest shape radius among the blocks) the stored inverse was built with
};

struct LpOff { int binv, xb, pi, w, b, f, ids, pos, rowbase, freebody, size; };

__host__ __device__ inline LpOff lp_layout(int MM, int MC) {
    LpOff o;
    int p = 0;
    o.binv = p; p += MM * MM * 8;
    o.xb = p; p += MM * 8;
    o.pi = p; p += MM * 8;
    o.w = p; p += MM * 8;
    o.b = p; p += MM * 8;
    o.f = p; p += 2 * MC * 8;
    o.ids = p; p += ((MM * 2 + 15) & ~15);
    o.pos = p; p += ((2 * MC + 15) & ~15);
    o.rowbase = p; p += ((NBODY + 15) & ~15);
    o.freebody = p; p += ((NB + 15) & ~15);
    o.size = (p + 15) & ~15;
    return o;
}

// monotone map float -> uint32 (a < b  <=>  key(a) < key(b))
__device__ __forceinline__ unsigned ordered_key(float x) {
    const unsigned u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct Lp {
    static constexpr double PIV_TOL = 1e-7;     // smallest pivot element
    static constexpr double HARRIS = 1e-9;      // feasibility slack of the ratio test
    static constexpr double D_TOL = 1e-10;      // reduced costs above -D_TOL count as non-negative
    static constexpr double CERT_REL = 1e-5;    // certificate of infeasibility: pi . r <= CERT_REL pi . b for EVERY ray

    // contacts of the environment (shared with the Newton solver)
    const double *G;
    const uint8_t *c_a, *c_b, *adj_ptr, *adj;
    // this problem
    double *Binv, *xB, *pi, *w, *b, *f;
    uint16_t *ids;            // basic column of every row position: LP_ART or ray = 2 * contact + sign
    uint8_t *pos;             // ray -> position in the basis, 0xFF = non-basic
    int8_t *rowbase;
    uint8_t *freebody;
    int MS;                   // row stride of Binv (= 3 * max_blocks, also the layout in HBM)
    int m, nfree, nc, lane, pivots;
    int why;                  // why the last run() returned LP_NONE (statistics: bw_debug_lp_stats)
    double mu, nb;
    unsigned long long artmask;   // positions that hold an artificial column
#ifdef BW_PROFILE
    long long tp[4];              // cycles: duals + pricing, column + ratio test, update, certificates
#define LP_T0(name) const long long name = clock64()
#define LP_ACC(i, t0) tp[i] += clock64() - (t0)
#else
#define LP_T0(name)
#define LP_ACC(i, t0)
#endif

    __device__ __forceinline__ double art_sum() const {
        double z = 0.0;
#pragma unroll 1
        for (int i = lane; i < m; i += 32)
            if ((artmask >> i) & 1ull) z += xB[i];
        return warp_sum(z);
    }

    // rows of the stored basis this problem can start from: the stored free blocks must be the first free blocks
    // of the problem (same blocks, new ones only above them), else 0
    __device__ static __forceinline__ int usable_rows(LpMeta meta, uint32_t free_mask) {
        const uint32_t mo = meta.mask;
        const int m_old = meta.m;
        if (m_old == 0 || (mo & ~free_mask) != 0u || m_old != 3 * __popc(mo) || m_old > 3 * __popc(free_mask)) return 0;
        const int hi = 32 - __clz(mo);
        const uint32_t below = (hi >= 32) ? 0xffffffffu : ((1u << hi) - 1u);
        return (((free_mask & ~mo) & below) == 0u) ? m_old : 0;
    }

    // the stored rows as they lie in HBM (row stride MS) -> shared memory; all threads of the CTA, 16-byte copies
    // (both addresses are 16-byte aligned: bw_create pads the per-environment stride to an even count)
    __device__ static __forceinline__ void load_rows(double *Binv, const double *gB, int ndoubles, int tid, int nthreads) {
        const int n2 = ndoubles >> 1;
        const double2 *src = reinterpret_cast<const double2 *>(gB);
        double2 *dst = reinterpret_cast<double2 *>(Binv);
#pragma unroll 8
        for (int q = tid; q < n2; q += nthreads) dst[q] = __ldcg(src + q);
        if ((ndoubles & 1) && tid == 0) Binv[ndoubles - 1] = gB[ndoubles - 1];
    }
    __device__ static __forceinline__ void store_rows(double *gB, const double *Binv, int ndoubles, int tid, int nthreads) {
        const int n2 = ndoubles >> 1;
        const double2 *src = reinterpret_cast<const double2 *>(Binv);
        double2 *dst = reinterpret_cast<double2 *>(gB);
#pragma unroll 4
        for (int q = tid; q < n2; q += nthreads) dst[q] = src[q];
        if ((ndoubles & 1) && tid == 0) gB[ndoubles - 1] = Binv[ndoubles - 1];
    }

    // Rows, right-hand side, and the stored basis of the environment extended to the free blocks `free_mask`
    // (rows of blocks the stored basis does not know start with their artificial columns).  Returns false when a
    // stored column no longer exists (never along real steps: the caller then drops the basis).
    // L0: the torque-row scale of this step's contact data.  A block with a larger radius changes it (mixed
    // libraries): the torque rows of every ray column are then s = L0_old / L0 times what the stored inverse was built
    // with, B' = D B E with D = diag(1, 1, s, ...) and E = 1/s on the artificial columns of torque rows (they are unit
    // vectors in either scaling), so B'^-1 = E^-1 B^-1 D^-1: torque COLUMNS of the inverse times 1/s, the ROWS of
    // torque-row artificials times s.
    __device__ bool setup(uint32_t free_mask, int n, const double *s_body, LpMeta meta, const uint16_t *gI,
                          const uint8_t *pair_itf, double L0) {
        const int m_old = usable_rows(meta, free_mask);
        // stored column identities: issued first, needed last
        const uint16_t gid0 = (lane < m_old) ? gI[lane] : LP_ART, gid1 = (lane + 32 < m_old) ? gI[lane + 32] : LP_ART;
        const bool is_free = lane < n && ((free_mask >> lane) & 1u);
        const unsigned fb = __ballot_sync(FULL, is_free);
        nfree = __popc(fb);
        m = 3 * nfree;
        const int myrow = __popc(fb & ((1u << lane) - 1));
        if (lane == 0) rowbase[0] = -1;
        if (lane < n) rowbase[lane + 1] = is_free ? (int8_t)(3 * myrow) : (int8_t)-1;
        if (is_free) freebody[myrow] = (uint8_t)(lane + 1);
        const double wgt = is_free ? s_body[(lane + 1) * 8 + 2] : 0.0;
        nb = sqrt(warp_sum(wgt * wgt));
        if (is_free) {
            b[3 * myrow] = 0.0;
            b[3 * myrow + 1] = wgt / nb;
            b[3 * myrow + 2] = 0.0;
        }
#pragma unroll 1
        for (int c = lane; c < 2 * nc; c += 32) pos[c] = 0xFF;
        if (m_old > 0 && meta.L0 != L0) {
            const double s = meta.L0 / L0, is = L0 / meta.L0;
#pragma unroll 1
            for (int i = 0; i < m_old; i++) {
                const bool art_torque = (i % 3 == 2) && (__shfl_sync(FULL, (i < 32) ? gid0 : gid1, i & 31) & LP_ART);
#pragma unroll 1
                for (int k = lane; k < m_old; k += 32) {
                    double v = Binv[i * MS + k];
                    if (k % 3 == 2) v *= is;
                    if (art_torque) v *= s;
                    Binv[i * MS + k] = v;
                }
            }
        }
        // basis inverse: the stored rows were copied by load_rows(); the identity for the new rows
        {
            const int dm = m - m_old;
            if (dm > 0) {
#pragma unroll 1
                for (int i = lane; i < m_old; i += 32)
                    for (int k = m_old; k < m; k++) Binv[i * MS + k] = 0.0;
#pragma unroll 1
                for (int i = m_old; i < m; i++)
                    for (int k = lane; k < m; k += 32) Binv[i * MS + k] = (i == k) ? 1.0 : 0.0;
            }
        }
        __syncwarp();
        bool bad = false;
        unsigned long long am = 0ull;
#pragma unroll 1
        for (int i0 = 0; i0 < m; i0 += 32) {
            const int i = i0 + lane;
            bool art = false;
            if (i < m) {
                uint16_t id = LP_ART;
                if (i < m_old) {
                    const uint16_t gid = i0 ? gid1 : gid0;
                    if (!(gid & LP_ART)) {
                        const int itf = pair_itf[gid >> 2];
                        if (itf == 0xFF || 2 * itf + 1 >= nc) bad = true;
                        else {
                            id = (uint16_t)(2 * (2 * itf + ((gid >> 1) & 1)) + (gid & 1));
                            pos[id] = (uint8_t)i;
                        }
                    }
                }
                ids[i] = id;
                art = (id & LP_ART) != 0;
                // basic solution of the new right-hand side: only the weight rows of b are non-zero
                double x;
                if (i < m_old) {
                    x = 0.0;
                    const double *Bi = Binv + i * MS;
#pragma unroll 2
                    for (int k = 1; k < m_old; k += 3) x = fma(Bi[k], b[k], x);
                    x = fmax(x, 0.0);
                } else {
                    x = b[i];
                }
                xB[i] = x;
            }
            am |= (unsigned long long)__ballot_sync(FULL, art) << i0;
        }
        artmask = am;
        pivots = 0;
        why = 6;
        __syncwarp();
        return !__any_sync(FULL, bad);
    }

    // ||b - A f|| of the basic solution (f = the forces the basic rays stand for), from the contact data
    __device__ double primal_residual() {
#pragma unroll 1
        for (int c = lane; c < nc; c += 32) {
            const int p0 = pos[2 * c], p1 = pos[2 * c + 1];
            const double lp = (p0 != 0xFF) ? xB[p0] : 0.0, lm = (p1 != 0xFF) ? xB[p1] : 0.0;
            f[2 * c] = lp + lm;
            f[2 * c + 1] = mu * (lp - lm);
        }
        __syncwarp();
        double acc = 0.0;
#pragma unroll 1
        for (int i = lane; i < m; i += 32) {
            const int I = i / 3, k = i - 3 * I;
            const int body = freebody[I];
            double af = 0.0;
#pragma unroll 2
            for (int q = adj_ptr[body]; q < adj_ptr[body + 1]; q++) {
                const int e = adj[q];
                const int c = e & 0x7f;
                const double *Gc = G + c * 12 + (e >> 7) * 6;
                af += Gc[k] * f[2 * c] + Gc[3 + k] * f[2 * c + 1];
            }
            const double r = b[i] - af;
            acc += r * r;
        }
        return sqrt(warp_sum(acc));
    }

    // Phase 1 from the basis left by setup().  r_exit: residual under which the system counts as feasible;
    // z_inf: a certificate of infeasibility needs pi . b above this (the optimal phase-1 objective z* bounds the
    // least-squares residual from below by z* / sqrt(m), so z_inf >= sqrt(m) stable_tol keeps the verdict rule).
    // Returns LP_FEASIBLE (res = ||b - A f|| <= r_exit), LP_INFEASIBLE (certificate), or LP_NONE.
    __device__ int run(double r_exit, double z_inf, double &res) {
        const int maxpiv = 4 * m + 16;
        why = 0;
#ifdef BW_PROFILE
        tp[0] = tp[1] = tp[2] = tp[3] = 0;
#endif
        double z = art_sum();                     // phase-1 objective: bounds ||b - A f|| of the basic solution
        res = nan("");
#pragma unroll 1
        while (true) {
            if (z <= r_exit) {
                z = art_sum();
                if (z <= r_exit) {
                    LP_T0(t_d);
                    const double r = primal_residual();
                    LP_ACC(3, t_d);
                    if (r <= r_exit) { res = r; return LP_FEASIBLE; }
                    why = 3;
                    return LP_NONE;
                }
            }
            if (pivots >= maxpiv) { why = 1; return LP_NONE; }
            LP_T0(t_a);
            // dual vector pi = sum of the artificial rows of the basis inverse (lane = column)
            {
                double p0 = 0.0, p1 = 0.0;
                const int k0 = lane, k1 = lane + 32;
#pragma unroll 1
                for (unsigned long long mm = artmask; mm; mm &= mm - 1) {
                    const int i = __ffsll((long long)mm) - 1;
                    if (k0 < m) p0 += Binv[i * MS + k0];
                    if (k1 < m) p1 += Binv[i * MS + k1];
                }
                if (k0 < m) pi[k0] = p0;
                if (k1 < m) pi[k1] = p1;
            }
            __syncwarp();
            // pricing: reduced cost of ray (c, +-) = -(pi . a_n +- mu pi . a_t); most negative non-basic one enters
            double best = 0.0, dall = 0.0;
            int bestray = -1;
#pragma unroll 1
            for (int c = lane; c < nc; c += 32) {
                const double *Gc = G + c * 12;
                const int ra = rowbase[c_a[c]], rb = rowbase[c_b[c]];
                double pn = 0.0, pt = 0.0;
                if (ra >= 0) {
                    const double v0 = pi[ra], v1 = pi[ra + 1], v2 = pi[ra + 2];
                    pn = Gc[0] * v0 + Gc[1] * v1 + Gc[2] * v2;
                    pt = Gc[3] * v0 + Gc[4] * v1 + Gc[5] * v2;
                }
                if (rb >= 0) {
                    const double v0 = pi[rb], v1 = pi[rb + 1], v2 = pi[rb + 2];
                    pn += Gc[6] * v0 + Gc[7] * v1 + Gc[8] * v2;
                    pt += Gc[9] * v0 + Gc[10] * v1 + Gc[11] * v2;
                }
                const double dp = -(pn + mu * pt), dm = -(pn - mu * pt);
                dall = fmin(dall, fmin(dp, dm));
                if (dp < best && pos[2 * c] == 0xFF) { best = dp; bestray = 2 * c; }
                if (dm < best && pos[2 * c + 1] == 0xFF) { best = dm; bestray = 2 * c + 1; }
            }
            const unsigned key = (bestray >= 0) ? ((ordered_key((float)best) & 0xffffff00u) | (unsigned)bestray) : 0xffffffffu;
            const unsigned kmin = __reduce_min_sync(FULL, key);
            double dq = 0.0;
            int q = -1;
            if (kmin != 0xffffffffu) {
                const unsigned own = __ballot_sync(FULL, key == kmin);
                const int src = __ffs(own) - 1;
                dq = __shfl_sync(FULL, best, src);
                q = (int)(kmin & 0xffu);
            }
            if (!(dq < -D_TOL)) {
                // optimal basis: pi is a Farkas vector if it clears every ray (basic ones included: their reduced
                // cost is zero only as far as the basis inverse is exact) and pi . b is clearly positive.  With
                // pi . r <= eps for all rays, any lambda >= 0 with R lambda = b has pi . b <= eps sum(lambda): the
                // certificate stands unless the contact forces add up to more than 1 / CERT_REL = 1e5 times the
                // weight of the structure (b is normalised)
                double dmin = dall;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(FULL, dmin, o));
                double pb = 0.0;
#pragma unroll 1
                for (int i = lane; i < m; i += 32) pb += pi[i] * b[i];
                pb = warp_sum(pb);
                if (pb > z_inf && dmin >= -CERT_REL * pb) return LP_INFEASIBLE;
                why = (pb > z_inf) ? 4 : 5;
                return LP_NONE;
            }
            LP_ACC(0, t_a);
            LP_T0(t_b);
            // entering column in the current basis: w = Binv r_q (six non-zeros)
            const int c = q >> 1;
            const double sg = (q & 1) ? -mu : mu;
            const double *Gc = G + c * 12;
            const int ra = rowbase[c_a[c]], rb = rowbase[c_b[c]];
            const double ca0 = Gc[0] + sg * Gc[3], ca1 = Gc[1] + sg * Gc[4], ca2 = Gc[2] + sg * Gc[5];
            const double cb0 = Gc[6] + sg * Gc[9], cb1 = Gc[7] + sg * Gc[10], cb2 = Gc[8] + sg * Gc[11];
            // Harris ratio test: bound from the relaxed ratios, then the largest pivot element under the bound
            unsigned krel = 0xffffffffu;
#pragma unroll 1
            for (int i = lane; i < m; i += 32) {
                const double *Bi = Binv + i * MS;
                double acc = 0.0;
                if (ra >= 0) acc = Bi[ra] * ca0 + Bi[ra + 1] * ca1 + Bi[ra + 2] * ca2;
                if (rb >= 0) acc += Bi[rb] * cb0 + Bi[rb + 1] * cb1 + Bi[rb + 2] * cb2;
                w[i] = acc;
                if (acc > PIV_TOL) krel = min(krel, __float_as_uint(__double2float_ru((xB[i] + HARRIS) * fast_rcp(acc) * (1.0 + 1e-9))));
            }
            krel = __reduce_min_sync(FULL, krel);
            if (krel == 0xffffffffu) { why = 2; return LP_NONE; }   // no positive pivot element: numerical trouble
            const double tmax = (double)__uint_as_float(krel);
            unsigned kpiv = 0u;
#pragma unroll 1
            for (int i = lane; i < m; i += 32) {
                const double wi = w[i];
                if (wi > PIV_TOL && xB[i] <= tmax * wi) kpiv = max(kpiv, (__float_as_uint((float)wi) & 0xffffffc0u) | (unsigned)i);
            }
            kpiv = __reduce_max_sync(FULL, kpiv);
            if (kpiv == 0u) { why = 2; return LP_NONE; }
            const int p = (int)(kpiv & 0x3fu);
            __syncwarp();
            LP_ACC(1, t_b);
            LP_T0(t_c);
            const double inv = 1.0 / w[p];
            const double theta = xB[p] * inv;
            // rank-one update of the basis inverse, lane = column; four rows per trip, loads first (the compiler cannot
            // move a load above a store that may alias it, and one row at a time is one shared-memory round trip per row)
            {
                const int k0 = lane, k1 = lane + 32;
                const bool h0 = k0 < m, h1 = k1 < m;
                const double t0 = h0 ? Binv[p * MS + k0] * inv : 0.0;
                const double t1 = h1 ? Binv[p * MS + k1] * inv : 0.0;
                double *c0 = Binv + (h0 ? k0 : 0), *c1 = Binv + (h1 ? k1 : 0);
#pragma unroll 1
                for (int i = 0; i < m; i += 4) {
                    // rows past the end and the pivot row itself take part with w = 0 (the value is written back as it is)
                    const int i0 = i, i1 = min(i + 1, m - 1), i2 = min(i + 2, m - 1), i3 = min(i + 3, m - 1);
                    const double w0 = (i0 == p) ? 0.0 : w[i0];
                    const double w1 = (i + 1 >= m || i1 == p) ? 0.0 : w[i1];
                    const double w2 = (i + 2 >= m || i2 == p) ? 0.0 : w[i2];
                    const double w3 = (i + 3 >= m || i3 == p) ? 0.0 : w[i3];
                    if ((w0 == 0.0) & (w1 == 0.0) & (w2 == 0.0) & (w3 == 0.0)) continue;       // uniform
                    const double a0 = c0[i0 * MS], a1 = c0[i1 * MS], a2 = c0[i2 * MS], a3 = c0[i3 * MS];
                    if (h1) {
                        const double e0 = c1[i0 * MS], e1 = c1[i1 * MS], e2 = c1[i2 * MS], e3 = c1[i3 * MS];
                        c1[i0 * MS] = fma(-w0, t1, e0);
                        if (i + 1 < m) c1[i1 * MS] = fma(-w1, t1, e1);
                        if (i + 2 < m) c1[i2 * MS] = fma(-w2, t1, e2);
                        if (i + 3 < m) c1[i3 * MS] = fma(-w3, t1, e3);
                    }
                    if (h0) {
                        c0[i0 * MS] = fma(-w0, t0, a0);
                        if (i + 1 < m) c0[i1 * MS] = fma(-w1, t0, a1);
                        if (i + 2 < m) c0[i2 * MS] = fma(-w2, t0, a2);
                        if (i + 3 < m) c0[i3 * MS] = fma(-w3, t0, a3);
                    }
                }
                if (h0) Binv[p * MS + k0] = t0;
                if (h1) Binv[p * MS + k1] = t1;
            }
#pragma unroll 1
            for (int i = lane; i < m; i += 32) xB[i] = (i == p) ? theta : fmax(fma(-w[i], theta, xB[i]), 0.0);
            if (lane == 0) {
                const uint16_t out = ids[p];
                if (!(out & LP_ART)) pos[out] = 0xFF;
                ids[p] = (uint16_t)q;
                pos[q] = (uint8_t)p;
            }
            artmask &= ~(1ull << p);
            z += theta * dq;
            pivots++;
            __syncwarp();
            LP_ACC(2, t_c);
        }
    }

    // keep the final basis for the next step (the matrix itself: store_rows)
    __device__ void store(LpMeta *gmeta, uint16_t *gI, uint32_t free_mask, const uint8_t *itf_pair, bool feasible,
                          double L0) const {
#pragma unroll 1
        for (int i = lane; i < m; i += 32) {
            const uint16_t id = ids[i];
            uint16_t g = LP_ART;
            if (!(id & LP_ART)) {
                const int c = id >> 1;
                g = (uint16_t)(((unsigned)itf_pair[c >> 1] << 2) | ((unsigned)(c & 1) << 1) | (unsigned)(id & 1));
            }
            gI[i] = g;
        }
        if (lane == 0) {
            LpMeta mt;
            mt.mask = free_mask;
            mt.m = (uint16_t)m;
            mt.feasible = feasible ? 1 : 0;
            mt.L0 = L0;
            *gmeta = mt;
        }
    }
};

}  // namespace bw

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

struct LpOff { int binv, xb, pi, w, b, ids, pos, crow, rowbase, freebody, size; };

// row stride of the basis inverse in shared memory and HBM: even, so that two columns are one 16-byte access
__host__ __device__ inline int lp_row_stride(int MM) { return (MM + 1) & ~1; }

// One problem = a set of vectors (`vec` bytes, offsets below) + its basis inverse (rows x lp_row_stride doubles).
// The region holds  [vectors 0 | vectors 1 | matrix 0 | matrix 1]: one problem at a time may use all 3 max_blocks rows
// (matrix 0 then runs over matrix 1), two problems side by side need (m0 + m1) rows to fit.
__host__ __device__ inline LpOff lp_layout(int MM, int MC, bool two_sets = true) {
    LpOff o;
    int p = 0;
    auto a16 = [](int x) { return (x + 15) & ~15; };
    o.xb = p; p += a16(MM * 8);
    o.pi = p; p += a16(MM * 8);
    o.w = p; p += a16((MM + 4) * 8);          // padded: the update reads w four rows at a time
    o.b = p; p += a16(MM * 8);
    o.ids = p; p += a16(MM * 2);
    o.pos = p; p += a16(2 * MC);
    o.crow = p; p += a16(2 * MC);
    o.rowbase = p; p += a16(NBODY);
    o.freebody = p; p += a16(NB);
    o.size = p;                               // bytes of one vector set
    o.binv = two_sets ? 2 * p : p;            // offset of matrix 0 (a small region holds one vector set only)
    return o;
}
// bytes the region needs for one problem with all rows
__host__ __device__ inline int lp_region_bytes(int MM, int MC, bool two_sets) {
    return lp_layout(MM, MC, two_sets).binv + ((MM + 3) & ~3) * lp_row_stride(MM) * 8;     // rows in fours (update)
}

// monotone map float -> uint32 (a < b  <=>  key(a) < key(b))
__device__ __forceinline__ unsigned ordered_key(float x) {
    const unsigned u = __float_as_uint(x);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

struct Lp {
    static constexpr double PIV_TOL = 1e-7;     // smallest pivot element
    static constexpr double HARRIS = 1e-9;      // feasibility slack of the ratio test
    static constexpr double D_TOL = 1e-9;       // reduced costs above -D_TOL count as non-negative
    static constexpr double CERT_REL = 1e-5;    // certificate of infeasibility: pi . r <= CERT_REL pi . b for EVERY ray

    // contacts of the environment (shared with the Newton solver)
    const double *G;
    const uint8_t *c_a, *c_b, *adj_ptr, *adj;
    // this problem
    double *Binv, *xB, *pi, *w, *b;
    uint16_t *ids;            // basic column of every row position: LP_ART or ray = 2 * contact + sign
    uint8_t *pos;             // ray -> position in the basis, 0xFF = non-basic
    uint16_t *crow;           // contact point -> (first row of body a + 1) | (first row of body b + 1) << 8, 0 = support
    int8_t *rowbase;
    uint8_t *freebody;
    int MS;                   // row stride of Binv (lp_row_stride(3 * max_blocks), also the layout in HBM)
    int m, nfree, nc, lane, pivots;
    double flops;             // work estimate of the run (bw_step_out.solver_kflops): per pivot 28 nc + 2 m^2 + 16 m
    int ndegen;               // pivots of the last run() that did not move (theta = 0): statistics
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
    // nb_: norm of the weights of the free blocks (b is normalised); gX: the stored basic solution in physical units.
    // gid0/1, gx0/1: the stored column identities and basic solution of rows lane and lane + 32 (loaded by the caller
    // long before they are needed).
    __device__ bool setup(uint32_t free_mask, int n, const double *s_body, LpMeta meta, uint16_t gid0, uint16_t gid1,
                          double gx0, double gx1, const uint8_t *pair_itf, double L0, double nb_) {
        const int m_old = usable_rows(meta, free_mask);
        if (lane >= m_old) { gid0 = LP_ART; gx0 = 0.0; }
        if (lane + 32 >= m_old) { gid1 = LP_ART; gx1 = 0.0; }
        const bool is_free = lane < n && ((free_mask >> lane) & 1u);
        const unsigned fb = __ballot_sync(FULL, is_free);
        nfree = __popc(fb);
        m = 3 * nfree;
        const int myrow = __popc(fb & ((1u << lane) - 1));
        if (lane == 0) rowbase[0] = -1;
        if (lane < n) rowbase[lane + 1] = is_free ? (int8_t)(3 * myrow) : (int8_t)-1;
        if (is_free) freebody[myrow] = (uint8_t)(lane + 1);
        const double wgt = is_free ? s_body[(lane + 1) * 8 + 2] : 0.0;
        nb = nb_;
        const double inv_nb = 1.0 / nb_;
        if (is_free) {
            b[3 * myrow] = 0.0;
            b[3 * myrow + 1] = wgt * inv_nb;
            b[3 * myrow + 2] = 0.0;
        }
        __syncwarp();
#pragma unroll 1
        for (int c = lane; c < nc; c += 32) {
            reinterpret_cast<uint16_t *>(pos)[c] = 0xFFFF;
            crow[c] = (uint16_t)((rowbase[c_a[c]] + 1) | ((rowbase[c_b[c]] + 1) << 8));
        }
        if (m_old > 0 && meta.L0 != L0) {
            const double s = meta.L0 / L0, is = L0 / meta.L0;
            // (the basic solution: D^-1 b = b because the torque rows of b are zero, so x' = E^-1 x)
            if ((lane % 3 == 2) && (gid0 & LP_ART)) gx0 *= s;
            if (((lane + 32) % 3 == 2) && (gid1 & LP_ART)) gx1 *= s;
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
                // basic solution: the stored one in the normalisation of this problem (the right-hand side of the old
                // rows is the old one up to that factor); new rows start with their artificial column at b
                const double x = (i < m_old) ? fmax((i0 ? gx1 : gx0) * inv_nb, 0.0) : b[i];
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

    // ||b - A f|| of the basic solution (f = the forces the basic rays stand for: fn = lambda+ + lambda-,
    // ft = mu (lambda+ - lambda-)), from the contact data
    __device__ double primal_residual() {
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
                const unsigned bp = reinterpret_cast<const uint16_t *>(pos)[c];
                const double lp = ((bp & 0xffu) != 0xffu) ? xB[bp & 0xffu] : 0.0, lm = ((bp >> 8) != 0xffu) ? xB[bp >> 8] : 0.0;
                const double *Gc = G + c * 12 + (e >> 7) * 6;
                af += Gc[k] * (lp + lm) + Gc[3 + k] * (mu * (lp - lm));
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
        const int maxpiv = 2 * m + 24;            // (an empty basis needs about 1.5 m pivots; warm runs 3, 30 at the most)
        why = 0;
        ndegen = 0;
        flops = 0.0;
#ifdef BW_PROFILE
        tp[0] = tp[1] = tp[2] = tp[3] = 0;
#endif
        double z = art_sum();                     // phase-1 objective: bounds ||b - A f|| of the basic solution
        res = nan("");
        // lane = two adjacent columns of the basis inverse (one 16-byte access): update and dual vector
        const int k2 = 2 * lane;
        const bool hk = k2 < m;
        const int MS2 = MS >> 1;                  // row stride in double2
        double2 *__restrict__ col = reinterpret_cast<double2 *>(Binv + (hk ? k2 : 0));
        double2 *pi2s = reinterpret_cast<double2 *>(pi);
        const int i0 = lane, i1 = lane + 32;      // lane = row(s) in the ratio test
        double2 pi2 = make_double2(0.0, 0.0);
        bool fresh = false;                       // pi comes straight from the basis inverse (not from updates)
        bool refreshed = false;                   // the basic solution was recomputed from the inverse
        bool blocked = false;                     // the last entering column had no pivot row
#pragma unroll 1
        while (true) {
            if (z <= r_exit) {
                z = art_sum();
                if (z <= r_exit) {
                    LP_T0(t_d);
                    const double r = primal_residual();
                    LP_ACC(3, t_d);
                    if (r <= r_exit) { res = r; return LP_FEASIBLE; }
                    // the updated basic solution has drifted from the basis: take it from the inverse again (once)
                    if (refreshed) { why = 3; return LP_NONE; }
                    refreshed = true;
                    __syncwarp();
#pragma unroll 1
                    for (int i = lane; i < m; i += 32) {
                        const double *Bi = Binv + i * MS;
                        double x = 0.0;
#pragma unroll 2
                        for (int k = 1; k < m; k += 3) x = fma(Bi[k], b[k], x);
                        xB[i] = fmax(x, 0.0);
                    }
                    __syncwarp();
                    z = art_sum();
                    if (z <= r_exit) {
                        const double r2 = primal_residual();
                        if (r2 <= r_exit) { res = r2; return LP_FEASIBLE; }
                        why = 3;
                        return LP_NONE;
                    }
                }
            }
            if (pivots >= maxpiv) { why = 1; return LP_NONE; }
            LP_T0(t_a);
            if (pivots == 0) {
                // dual vector pi = sum of the artificial rows of the basis inverse; kept up to date by the pivots below
                // (pi += d_q x new pivot row) and recomputed before an optimal basis is believed
                pi2 = make_double2(0.0, 0.0);
#pragma unroll 1
                for (unsigned long long mm = artmask; mm; mm &= mm - 1) {
                    const int i = __ffsll((long long)mm) - 1;
                    const double2 v = col[i * MS2];
                    pi2.x += v.x; pi2.y += v.y;
                }
                if (hk) pi2s[lane] = pi2;
                fresh = true;
                __syncwarp();
            }
            // pricing: reduced cost of ray (c, +-) = -(pi . a_n +- mu pi . a_t); most negative non-basic one enters
            double best = 0.0, dall = 0.0;
            int bestray = -1;
#pragma unroll 2
            for (int c = lane; c < nc; c += 32) {
                const double2 *G2 = reinterpret_cast<const double2 *>(G + c * 12);
                const unsigned cr = crow[c];
                const int ra = (int)(cr & 0xffu) - 1, rb = (int)(cr >> 8) - 1;
                const unsigned bp = reinterpret_cast<const uint16_t *>(pos)[c];     // pos[2c] | pos[2c + 1] << 8
                double pn = 0.0, pt = 0.0;
                if (ra >= 0) {
                    const double2 g0 = G2[0], g1 = G2[1], g2 = G2[2];               // n0 n1 | n2 t0 | t1 t2
                    const double v0 = pi[ra], v1 = pi[ra + 1], v2 = pi[ra + 2];
                    pn = g0.x * v0 + g0.y * v1 + g1.x * v2;
                    pt = g1.y * v0 + g2.x * v1 + g2.y * v2;
                }
                if (rb >= 0) {
                    const double2 g3 = G2[3], g4 = G2[4], g5 = G2[5];
                    const double v0 = pi[rb], v1 = pi[rb + 1], v2 = pi[rb + 2];
                    pn += g3.x * v0 + g3.y * v1 + g4.x * v2;
                    pt += g4.y * v0 + g5.x * v1 + g5.y * v2;
                }
                const double dp = -(pn + mu * pt), dm = -(pn - mu * pt);
                dall = fmin(dall, fmin(dp, dm));
                if (dp < best && (bp & 0xffu) == 0xffu) { best = dp; bestray = 2 * c; }
                if (dm < best && (bp >> 8) == 0xffu) { best = dm; bestray = 2 * c + 1; }
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
            const bool optimal = !(dq < -D_TOL) || blocked;
            // (an attempt is also made as soon as the most negative reduced cost is small against the objective: what is
            // left to gain is then rounding noise of the duals, not a direction of descent)
            if (optimal || (z > z_inf && dq >= -CERT_REL * z)) {
                // optimal basis: pi is a Farkas vector if it clears every ray (basic ones included: their reduced
                // cost is zero only as far as the basis inverse is exact) and pi . b is clearly positive.  With
                // pi . r <= eps for all rays, any lambda >= 0 with R lambda = b has pi . b <= eps sum(lambda): the
                // certificate stands unless the contact forces add up to more than 1 / CERT_REL = 1e5 times the
                // weight of the structure (b is normalised).  The test holds for ANY vector pi, so the updated duals
                // are tried as they are; only when they fail it are they recomputed from the basis inverse and the
                // rays priced again (the run then either goes on or ends with the fresh duals' verdict).
                double dmin = dall;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) dmin = fmin(dmin, __shfl_xor_sync(FULL, dmin, o));
                double pb = 0.0;
#pragma unroll 1
                for (int i = lane; i < m; i += 32) pb += pi[i] * b[i];
                pb = warp_sum(pb);
                if (pb > z_inf && dmin >= -CERT_REL * pb) return LP_INFEASIBLE;
                if (!fresh) {
                    pi2 = make_double2(0.0, 0.0);
#pragma unroll 1
                    for (unsigned long long mm = artmask; mm; mm &= mm - 1) {
                        const int i = __ffsll((long long)mm) - 1;
                        const double2 v = col[i * MS2];
                        pi2.x += v.x; pi2.y += v.y;
                    }
                    __syncwarp();
                    if (hk) pi2s[lane] = pi2;
                    fresh = true;
                    __syncwarp();
                    continue;
                }
                if (optimal) {
                    why = (pb > z_inf) ? 4 : 5;
                    return LP_NONE;
                }
            }
            LP_ACC(0, t_a);
            LP_T0(t_b);
            // entering column in the current basis: w = Binv r_q (six non-zeros), lane = row (two above 32 rows)
            const int c = q >> 1;
            const double sg = (q & 1) ? -mu : mu;
            const double2 *G2 = reinterpret_cast<const double2 *>(G + c * 12);
            const int ra = (int)(crow[c] & 0xffu) - 1, rb = (int)(crow[c] >> 8) - 1;
            double a0 = 0.0, a1 = 0.0, x0 = 0.0, x1 = 0.0;
            {
                const double2 g0 = G2[0], g1 = G2[1], g2 = G2[2], g3 = G2[3], g4 = G2[4], g5 = G2[5];
                const double ca0 = g0.x + sg * g1.y, ca1 = g0.y + sg * g2.x, ca2 = g1.x + sg * g2.y;
                const double cb0 = g3.x + sg * g4.y, cb1 = g3.y + sg * g5.x, cb2 = g4.x + sg * g5.y;
                if (i0 < m) {
                    const double *Bi = Binv + i0 * MS;
                    if (ra >= 0) a0 = Bi[ra] * ca0 + Bi[ra + 1] * ca1 + Bi[ra + 2] * ca2;
                    if (rb >= 0) a0 += Bi[rb] * cb0 + Bi[rb + 1] * cb1 + Bi[rb + 2] * cb2;
                    x0 = xB[i0];
                }
                if (i1 < m) {
                    const double *Bi = Binv + i1 * MS;
                    if (ra >= 0) a1 = Bi[ra] * ca0 + Bi[ra + 1] * ca1 + Bi[ra + 2] * ca2;
                    if (rb >= 0) a1 += Bi[rb] * cb0 + Bi[rb + 1] * cb1 + Bi[rb + 2] * cb2;
                    x1 = xB[i1];
                }
            }
            // Harris ratio test: bound from the relaxed ratios, then the largest pivot element under the bound
            unsigned krel = 0xffffffffu;
            if (a0 > PIV_TOL) krel = __float_as_uint(__double2float_ru((x0 + HARRIS) * fast_rcp(a0) * (1.0 + 1e-9)));
            if (a1 > PIV_TOL) krel = min(krel, __float_as_uint(__double2float_ru((x1 + HARRIS) * fast_rcp(a1) * (1.0 + 1e-9))));
            krel = __reduce_min_sync(FULL, krel);
            if (krel == 0xffffffffu) {
                // no positive pivot element: the column cannot enter.  With a reduced cost that is only noise this is an
                // optimal basis in disguise: let the certificate decide (it looks at every ray with fresh duals)
                if (blocked || dq < -1e-6) { why = 2; return LP_NONE; }
                blocked = true;
                fresh = false;
                continue;
            }
            const double tmax = (double)__uint_as_float(krel);
            unsigned kpiv = 0u;
            if (a0 > PIV_TOL && x0 <= tmax * a0) kpiv = (__float_as_uint((float)a0) & 0xffffffc0u) | (unsigned)i0;
            if (a1 > PIV_TOL && x1 <= tmax * a1) kpiv = max(kpiv, (__float_as_uint((float)a1) & 0xffffffc0u) | (unsigned)i1);
            kpiv = __reduce_max_sync(FULL, kpiv);
            if (kpiv == 0u) { why = 2; return LP_NONE; }
            const int p = (int)(kpiv & 0x3fu);
            const double wp = __shfl_sync(FULL, (p < 32) ? a0 : a1, p & 31);
            const double xp = __shfl_sync(FULL, (p < 32) ? x0 : x1, p & 31);
            double inv;
            {   // 1 / wp: hardware approximation + two Newton steps (full precision for a normal wp > PIV_TOL)
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(inv) : "d"(wp));
                inv = fma(fma(-wp, inv, 1.0), inv, inv);
                inv = fma(fma(-wp, inv, 1.0), inv, inv);
            }
            const double theta = xp * inv;
            // entering column for the update (the pivot row takes part as a no-op: w = 0; rows past the end likewise,
            // the update reads four rows at a time) and the new basic solution
            if (i0 < m) { w[i0] = (i0 == p) ? 0.0 : a0; xB[i0] = (i0 == p) ? theta : fmax(fma(-a0, theta, x0), 0.0); }
            else if (i0 < m + 4) w[i0] = 0.0;
            if (i1 < m) { w[i1] = (i1 == p) ? 0.0 : a1; xB[i1] = (i1 == p) ? theta : fmax(fma(-a1, theta, x1), 0.0); }
            else if (i1 < m + 4) w[i1] = 0.0;
            __syncwarp();
            LP_ACC(1, t_b);
            LP_T0(t_c);
            // rank-one update of the basis inverse: four rows per trip with all loads in front of the stores (the
            // compiler cannot move a load above a store that may alias it, and a row at a time is a shared-memory
            // round trip per row); rows with w = 0 are skipped
            {
                double2 t = col[p * MS2];
                t.x *= inv; t.y *= inv;
                const double2 *__restrict__ w2 = reinterpret_cast<const double2 *>(w);
#pragma unroll 1
                for (int i = 0; i < m; i += 4) {
                    const double2 wa = w2[i >> 1], wb = w2[(i >> 1) + 1];
                    if ((wa.x == 0.0) & (wa.y == 0.0) & (wb.x == 0.0) & (wb.y == 0.0)) continue;     // uniform
                    // (the matrix is allocated in fours of rows: the rows past the end carry w = 0 and are written
                    // back as they are)
                    double2 *r = col + i * MS2;
                    const double2 b0 = r[0], b1 = r[MS2], b2 = r[2 * MS2], b3 = r[3 * MS2];
                    if (hk) {
                        r[0] = make_double2(fma(-wa.x, t.x, b0.x), fma(-wa.x, t.y, b0.y));
                        r[MS2] = make_double2(fma(-wa.y, t.x, b1.x), fma(-wa.y, t.y, b1.y));
                        r[2 * MS2] = make_double2(fma(-wb.x, t.x, b2.x), fma(-wb.x, t.y, b2.y));
                        r[3 * MS2] = make_double2(fma(-wb.y, t.x, b3.x), fma(-wb.y, t.y, b3.y));
                    }
                }
                if (hk) col[p * MS2] = t;
                // duals of the new basis: pi += d_q x (new pivot row)
                pi2.x = fma(dq, t.x, pi2.x);
                pi2.y = fma(dq, t.y, pi2.y);
                if (hk) pi2s[lane] = pi2;
                fresh = false;
            }
            if (lane == 0) {
                const uint16_t out = ids[p];
                if (!(out & LP_ART)) pos[out] = 0xFF;
                ids[p] = (uint16_t)q;
                pos[q] = (uint8_t)p;
            }
            artmask &= ~(1ull << p);
            z += theta * dq;
            ndegen += (theta <= 1e-13) ? 1 : 0;
            pivots++;
            flops += 28.0 * nc + 2.0 * m * m + 16.0 * m;
            __syncwarp();
            LP_ACC(2, t_c);
        }
    }

    // keep the final basis for the next step (the matrix itself: store_rows)
    __device__ void store(LpMeta *gmeta, uint16_t *gI, double *gX, uint32_t free_mask, const uint8_t *itf_pair,
                          bool feasible, double L0) const {
#pragma unroll 1
        for (int i = lane; i < m; i += 32) {
            gX[i] = xB[i] * nb;
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

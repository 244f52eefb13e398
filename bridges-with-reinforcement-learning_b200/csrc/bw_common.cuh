// Shared device-side definitions of the bridges_b200 kernels.
//
// Canonical arithmetic: every geometric quantity that decides a
// raster bit, a bounds flag or a target hit is computed in float64 with individually
// rounded multiplies and adds -- written with the __d*_rn intrinsics so that neither
// nvcc's FMA contraction nor -use_fast_math can change a bit.  The solver is free to
// use FMAs (its results are compared with a tolerance).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/bridges_b200.h"

namespace bw {

constexpr int NB = BW_MAX_BLOCKS;          // 16
constexpr int NBODY = NB + 1;              // floor + blocks; body 0 = floor, body j+1 = block j
constexpr int NF = BW_MAX_FACES;           // 6
constexpr int NV = BW_MAX_VERTS;           // 6
constexpr int IMG = BW_IMG;                // 64
constexpr int MAXITF = BW_MAX_INTERFACES;  // 48
constexpr int MAXC = 2 * MAXITF;           // contact points
constexpr int ORDER_KEYS = 32;             // cost classes of the step kernel's CTA order

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dadd_rn(a, -b); }

// R(x, z) = (c*x + s*z, c*z - s*x): rotation about +y, canonical operation order
__device__ __forceinline__ void rot(double c, double s, double x, double z, double &ox, double &oz) {
    ox = dadd(dmul(c, x), dmul(s, z));
    oz = dsub(dmul(c, z), dmul(s, x));
}

// block library entry as stored on the device
struct ShapeDev {
    int32_t n_faces, n_verts;
    uint32_t target_faces_mask, receiving_faces_mask;
    double face_nx[NF], face_nz[NF], face_cx[NF], face_cz[NF];
    double end0_x[NF], end0_z[NF], end1_x[NF], end1_z[NF];
    double vert_x[NV], vert_z[NV];
    double com_x, com_z, area, depth;
    double radius;   // max distance com -> vertex (torque row scale)
};

// per-environment task + bookkeeping (AssemblyGym attributes)
struct TaskDev {
    int32_t n_obstacles, n_targets;
    double obstacle_xz[BW_MAX_OBSTACLES][2];
    double target_xz[BW_MAX_TARGETS][2];
    int8_t remaining[BW_MAX_TARGETS];   // targets_remaining, as indices into targets, list order
    int8_t reached[BW_MAX_TARGETS];     // targets_reached, in the order they were reached
    int8_t n_remaining, n_reached;
    int8_t pad[6];
};

struct Pose {
    double x, z, c, s;
};

struct LpMeta;   // bw_lp.cuh

// everything a kernel needs about the handle (passed by value)
struct Params {
    int32_t E;
    int32_t max_steps;
    int32_t n_shapes;
    int32_t max_blocks;       // capacity used for shared-memory sizing (<= NB)
    int32_t max_itf;          // interface capacity used for shared-memory sizing (<= MAXITF)
    double xlim0, xlim1, ylim0, ylim1;
    double floor_halfwidth, floor_depth, density, tmax, amin, stable_tol;
    double inv_step_x, inv_step_y;       // 1/pixel pitch (only for conservative index ranges)
    const double *xs;          // [IMG] pixel x coordinates (numpy linspace semantics)
    const double *ys;          // [IMG] pixel z coordinates, row 0 = top
    const ShapeDev *shapes;    // [n_shapes]
    const ShapeDev *marker;    // the 0.6 cube of obstacles / target markers (cube06.urdf)
    int32_t collision_mode;    // 0: flags constant False; 1: polygon penetration (assembly_env.py:346-391)
    int32_t screen;            // 1: rigid-mechanism certificates before the equilibrium solve (bw_solver.cuh)
    double collision_tol;
    double bounds_lo[3], bounds_hi[3];
    // state
    int32_t *n_blocks;         // [E]
    Pose *pose;                // [E][NB]
    uint8_t *shape_of;         // [E][NB]
    uint8_t *face_occ;         // [E][NB] bit f = face f occupied (block_graph, gym_env.py:229-232)
    uint32_t *static_mask;     // [E] bit i = block i is a support
    uint64_t *block_bits;      // [E][IMG] raster of the placed blocks
    uint64_t *obst_bits;       // [E][IMG] raster of the obstacles
    float *reward_img;         // [E][IMG*IMG] Gaussian-blurred target raster
    TaskDev *task;             // [E]
    double *mu;                // [E]
    uint8_t *done;             // [E] last step returned terminated | truncated
    bw_step_out *last_out;     // [E] copy of the last step result (binary features of observe)
    uint8_t *su_valid;         // [E] last_out[e].stable_unfrozen describes the current blocks and supports
    // dual iterates of the last feasible solves, per block (physical units: y * ||weights||), the starting
    // points of the next step's solves (step_kernel phase 3)
    double *warm_y;            // [E][2][NB][3]: 0 = supports as step() leaves them, 1 = last block released
    uint8_t *warm_ok;          // [E][2] the entry holds the iterate of a solve that ended feasible
    int32_t warm_start;        // 0: every solve starts from y = 0 (tuning hook BW_NO_WARM)
    int32_t pad_ws;
    // CTA -> environment order of the step kernel: the environments queue up for the NEXT launch by expected
    // cost (key = 2 * blocks + "the frozen solve cannot be skipped", heaviest first), so that the long solves
    // start with the first wave of CTAs and the short ones fill in behind them.  Three queues in rotation:
    // launch t reads queue t % 3, fills queue (t + 1) % 3 and clears the counters of queue (t + 2) % 3.
    int32_t *order_cnt;        // [3][ORDER_KEYS]
    int32_t *order_q;          // [3][ORDER_KEYS][E]
    int32_t order_phase;       // queue read by the next launch (host side, advanced by launch_step)
    int32_t order_on;          // 0: CTA i = environment i (tuning hook BW_NO_ORDER)
    // shared-memory diet of multi-wave launches (more environments than CTA slots): the two problems of an
    // environment share one packed matrix and take turns (share_h), the block library and the pixel nodes stay in
    // global memory (lib_in_smem = 0) -- both buy resident CTAs per SM, which is what a latency-bound kernel needs
    // once there is more than one wave of work
    int32_t share_h;
    int32_t lib_in_smem;
    // warm-started LP verdicts of real steps (bw_lp.cuh): the optimal basis of the released problem of the last step
    int32_t lp_on;             // 0: every verdict by the screen + Newton path (tuning hook BW_NO_LP)
    int32_t lp_par;            // 1: frozen and released problem side by side on the two warps when their rows fit (BW_LP_SEQ: off)
    int32_t lp_pad;
    int32_t lp_stride;         // doubles per environment in lp_binv: (3 max_blocks)^2 rounded up to an even count
    LpMeta *lp_meta;           // [E]
    double *lp_binv;           // [E][lp_stride] basis inverse, row stride 3 max_blocks
    double *lp_xb;             // [E][3 NB] basic solution of the stored basis in physical units (times ||weights||)
    uint16_t *lp_ids;          // [E][3 NB] basic columns: LP_ART or pair index << 2 | contact point << 1 | ray sign
    unsigned long long *lp_stats;   // [32] counters of the LP path (tools/ only; see bw_debug_lp_stats)
    int32_t *cand_need;        // [1] largest untruncated candidate count an enumeration had to cut to `amax`
    int32_t *reset_err;        // [1] environments whose reset task was refused (bad shape index / too many blocks)
};

// ---------------------------------------------------------------- placement (K1)
// create_block (gym_env.py:204-216) -> align_frames_2d (geometry.py:39-50) in canonical
// arithmetic.  Returns 0 and the new pose, 1 for invalid indices, 2 when the env is full.
__device__ inline int place_block(const Params &P, const Pose *poses, const uint8_t *shape_of, int n,
                                  const bw_action &act, Pose &out) {
    bool ok = act.shape >= 0 && act.shape < P.n_shapes && act.face >= 0 && act.target_block >= -1 &&
              act.target_block < n;
    if (ok) ok = act.face < P.shapes[act.shape].n_faces;
    if (ok && act.target_block >= 0)
        ok = act.target_face >= 0 && act.target_face < P.shapes[shape_of[act.target_block]].n_faces;
    if (!ok) return 1;
    if (n >= P.max_blocks) return 2;
    double p1x = 0.0, p1z = 0.0, n1x = 0.0, n1z = 1.0;   // Frame.worldXY(): floor
    if (act.target_block >= 0) {
        const Pose tp = poses[act.target_block];
        const ShapeDev &ts = P.shapes[shape_of[act.target_block]];
        rot(tp.c, tp.s, ts.face_nx[act.target_face], ts.face_nz[act.target_face], n1x, n1z);
        double rx, rz;
        rot(tp.c, tp.s, ts.face_cx[act.target_face], ts.face_cz[act.target_face], rx, rz);
        p1x = dadd(rx, tp.x);
        p1z = dadd(rz, tp.z);
    }
    const ShapeDev &sh = P.shapes[act.shape];
    const double n2x = sh.face_nx[act.face], n2z = sh.face_nz[act.face];
    const double p2x = sh.face_cx[act.face], p2z = sh.face_cz[act.face];
    double c = -dadd(dmul(n1x, n2x), dmul(n1z, n2z));
    c = fmin(1.0, fmax(-1.0, c));
    const double cross_y = dsub(dmul(n1z, n2x), dmul(n1x, n2z));
    double s = fabs(cross_y);
    if (!(dadd(cross_y, 1e-6) > 0.0)) s = -s;      // the "+1e-6" axis of geometry.py:47
    const double wx = dadd(dadd(dmul(n1z, act.offset_x), dmul(n1x, act.offset_y)), p1x);
    const double wz = dadd(dsub(dmul(n1z, act.offset_y), dmul(n1x, act.offset_x)), p1z);
    double rx, rz;
    rot(c, s, p2x, p2z, rx, rz);
    out.x = dsub(wx, rx);
    out.z = dsub(wz, rz);
    out.c = c;
    out.s = s;
    return 0;
}

// posed bounding box of a shape
__device__ inline void posed_aabb(const ShapeDev &sh, const Pose &ps, double &xmin, double &xmax, double &zmin,
                                  double &zmax) {
    xmin = 1e300; xmax = -1e300; zmin = 1e300; zmax = -1e300;
    for (int v = 0; v < sh.n_verts; v++) {
        double vx, vz;
        rot(ps.c, ps.s, sh.vert_x[v], sh.vert_z[v], vx, vz);
        vx = dadd(vx, ps.x);
        vz = dadd(vz, ps.z);
        xmin = fmin(xmin, vx); xmax = fmax(xmax, vx);
        zmin = fmin(zmin, vz); zmax = fmax(zmax, vz);
    }
}

// ---------------------------------------------------------------- collision (K1, collision_mode = 1)
// _check_collision (assembly_env.py:346-391): "contact distance < -tol" as the exact penetration of
// two convex polygons.  sep = max over the faces of both polygons of the min over the other polygon's
// vertices of (v - c).n; collision iff sep < -tol, i.e. iff the min of EVERY face is below -tol.
// A FaceView is a posed face table: normals, centres and the two end points of every face (the end
// points of all faces are the polygon's vertices), `stride` doubles apart, translated by (tx, tz).
struct FaceView {
    const double *nx, *nz, *cx, *cz, *e0x, *e0z, *e1x, *e1z;
    int stride, nf;
    double tx, tz;
};

// true iff every face of `a` sees all vertices of `b` deeper than `tol` behind it
__device__ inline bool faces_all_penetrated(const FaceView &a, const FaceView &b, double tol) {
    for (int f = 0; f < a.nf; f++) {
        const int i = f * a.stride;
        const double nx = a.nx[i], nz = a.nz[i];
        const double cx = dadd(a.cx[i], a.tx), cz = dadd(a.cz[i], a.tz);
        double smin = INFINITY;
        for (int g = 0; g < b.nf; g++) {
            const int j = g * b.stride;
            const double v0 = dadd(dmul(dsub(dadd(b.e0x[j], b.tx), cx), nx), dmul(dsub(dadd(b.e0z[j], b.tz), cz), nz));
            const double v1 = dadd(dmul(dsub(dadd(b.e1x[j], b.tx), cx), nx), dmul(dsub(dadd(b.e1z[j], b.tz), cz), nz));
            smin = fmin(smin, fmin(v0, v1));
        }
        if (!(smin < -tol)) return false;      // a separating (or merely touching) axis
    }
    return true;
}

__device__ inline bool polygons_collide(const FaceView &a, const FaceView &b, double tol) {
    return faces_all_penetrated(a, b, tol) && faces_all_penetrated(b, a, tol);
}

__device__ inline FaceView face_view_posed(const double *faces8, int nf) {   // [nf][8] = nx nz cx cz e0x e0z e1x e1z
    FaceView v;
    v.nx = faces8; v.nz = faces8 + 1; v.cx = faces8 + 2; v.cz = faces8 + 3;
    v.e0x = faces8 + 4; v.e0z = faces8 + 5; v.e1x = faces8 + 6; v.e1z = faces8 + 7;
    v.stride = 8; v.nf = nf; v.tx = 0.0; v.tz = 0.0;
    return v;
}

__device__ inline FaceView face_view_shape(const ShapeDev &sh, double tx, double tz) {   // unrotated shape at (tx, tz)
    FaceView v;
    v.nx = sh.face_nx; v.nz = sh.face_nz; v.cx = sh.face_cx; v.cz = sh.face_cz;
    v.e0x = sh.end0_x; v.e0z = sh.end0_z; v.e1x = sh.end1_x; v.e1z = sh.end1_z;
    v.stride = 1; v.nf = sh.n_faces; v.tx = tx; v.tz = tz;
    return v;
}

// ---------------------------------------------------------------- raster (K4)
// Bits of image row `row` covered by shape `sh` posed at `ps`: contains_2d
// (assembly_env.py:126-137) at the pixel nodes of render_blocks_2d (rendering.py:105-113).
// A pixel is inside iff every half-plane value v = (px-cx)*nx + (pz-cz)*nz, evaluated with
// individually rounded operations, is <= 0.  Along a row v is weakly monotone in the column index
// (rounded subtraction, multiplication by a constant and addition all preserve order), so each face
// admits an interval of columns; its end is located by an estimate of the crossing and then fixed
// with the exact test, which makes the result identical to testing all 64 pixels one by one.
__device__ __forceinline__ bool pixel_in_halfplane(double px, double cx, double nx, double vz) {
    return dadd(dmul(dsub(px, cx), nx), vz) <= 0.0;
}

// posed half-planes + pixel window of one block: everything about a block the raster needs
struct PosedShape {
    double nx[NF], nz[NF], cx[NF], cz[NF], inv_nx[NF];
    int n_faces;
    int j_lo, j_hi, i_lo, i_hi;     // conservative window of pixel columns / rows (may be empty)
};

__device__ inline void pose_shape(const Params &P, const ShapeDev &sh, const Pose &ps, PosedShape &o) {
    double xmin, xmax, zmin, zmax;
    posed_aabb(sh, ps, xmin, xmax, zmin, zmax);
    o.j_lo = max((int)floor((xmin - P.xlim0) * P.inv_step_x) - 1, 0);
    o.j_hi = min((int)ceil((xmax - P.xlim0) * P.inv_step_x) + 1, IMG - 1);
    o.i_lo = max((int)floor((P.ylim1 - zmax) * P.inv_step_y) - 1, 0);
    o.i_hi = min((int)ceil((P.ylim1 - zmin) * P.inv_step_y) + 1, IMG - 1);
    o.n_faces = sh.n_faces;
    for (int k = 0; k < NF; k++) {
        if (k < sh.n_faces) {
            double ax, az;
            rot(ps.c, ps.s, sh.face_nx[k], sh.face_nz[k], o.nx[k], o.nz[k]);
            rot(ps.c, ps.s, sh.face_cx[k], sh.face_cz[k], ax, az);
            o.cx[k] = dadd(ax, ps.x);
            o.cz[k] = dadd(az, ps.z);
            o.inv_nx[k] = (o.nx[k] != 0.0) ? 1.0 / o.nx[k] : 0.0;
        }
    }
}

// bits of image row `row` for a posed shape given by arrays (which may live in shared memory)
// `stride` = distance (in doubles) between consecutive faces in nx / nz / cx / cz (inv_nx is dense).
__device__ inline uint64_t raster_row_posed(const Params &P, int n_faces, const double *nx, const double *nz,
                                            const double *cx, const double *cz, const double *inv_nx, int j_lo,
                                            int j_hi, int row, int stride = 1) {
    const double pz = P.ys[row];
    int lo = j_lo, hi = j_hi;                 // surviving column interval [lo, hi]
    for (int k = 0; k < n_faces && lo <= hi; k++) {
        const double fnx = nx[k * stride], fcx = cx[k * stride];
        const double vz = dmul(dsub(pz, cz[k * stride]), nz[k * stride]);
        if (fnx == 0.0) {                     // the value does not depend on the column
            if (!pixel_in_halfplane(P.xs[lo], fcx, fnx, vz)) hi = lo - 1;
            continue;
        }
        // estimated crossing column of v(px) = 0 (only an estimate: the exact test below decides)
        const double g = (fcx - vz * inv_nx[k] - P.xlim0) * P.inv_step_x;
        if (fnx > 0.0) {                      // v increases with the column: inside = [lo, b]
            int b = (g >= (double)hi) ? hi : (g < (double)lo ? lo - 1 : (int)floor(g));
            while (b < hi && pixel_in_halfplane(P.xs[b + 1], fcx, fnx, vz)) b++;
            while (b >= lo && !pixel_in_halfplane(P.xs[b], fcx, fnx, vz)) b--;
            hi = b;
        } else {                              // v decreases with the column: inside = [b, hi]
            int b = (g <= (double)lo) ? lo : (g > (double)hi ? hi + 1 : (int)ceil(g));
            while (b > lo && pixel_in_halfplane(P.xs[b - 1], fcx, fnx, vz)) b--;
            while (b <= hi && !pixel_in_halfplane(P.xs[b], fcx, fnx, vz)) b++;
            lo = b;
        }
    }
    if (lo > hi) return 0;
    const int w = hi - lo + 1;
    return (w >= 64) ? ~0ull : (((1ull << w) - 1) << lo);
}

// The same row mask for warps whose lanes hold rows of DIFFERENT posed shapes (enumerate_kernel): the two
// directions of a face (v increasing / decreasing with the column) walk one instruction sequence with a signed
// step instead of two divergent branches.  Same estimate, same exact tests, same result as raster_row_posed.
__device__ inline uint64_t raster_row_posed_mixed(const Params &P, int n_faces, const double *nx, const double *nz,
                                                  const double *cx, const double *cz, const double *inv_nx, int j_lo,
                                                  int j_hi, int row) {
    const double pz = P.ys[row];
    int lo = j_lo, hi = j_hi;                 // surviving column interval [lo, hi]
    for (int k = 0; k < n_faces && lo <= hi; k++) {
        const double fnx = nx[k], fcx = cx[k];
        const double vz = dmul(dsub(pz, cz[k]), nz[k]);
        if (fnx == 0.0) {                     // the value does not depend on the column
            if (!pixel_in_halfplane(P.xs[lo], fcx, fnx, vz)) hi = lo - 1;
            continue;
        }
        const double off = vz * inv_nx[k];    // cx - px* of the crossing
        const double g = (fcx - off - P.xlim0) * P.inv_step_x;
        const bool inc = fnx > 0.0;           // inside = [lo, b] (inc) or [b, hi] (!inc)
        const int step = inc ? 1 : -1;
        const int near = inc ? lo : hi, far = inc ? hi : lo;
        // b = estimated last inside column walking from `near` towards `far` (one beyond `near` = none)
        const double gr = inc ? floor(g) : ceil(g);
        int b = ((gr - (double)far) * (double)step >= 0.0) ? far
              : (((gr - (double)near) * (double)step < 0.0) ? near - step : (int)gr);
        // The exact test only has to decide when the estimate is close to a pixel node.  The rounded value
        // fl(fl(fl(px - cx) nx) + vz) has the sign of (px - cx) nx (1 + d) + vz with |d| < 2.3e-16, i.e. the test
        // moves the true crossing px* = cx - vz / nx by less than 2.3e-16 |px* - cx|; g carries a handful of
        // roundings of the same size (and the nodes P.xs are within an ulp of their ideal positions).  With
        // |px* - cx| < 1e3 all of that stays below 1e-11 pixels: an estimate further than 1e-9 from every node is
        // on the same side of it as the exact crossing.
        if (!(fabs(off) < 1e3 && fabs(g - rint(g)) > 1e-9)) {
            while (b != far && pixel_in_halfplane(P.xs[b + step], fcx, fnx, vz)) b += step;
            while ((b - near) * step >= 0 && !pixel_in_halfplane(P.xs[b], fcx, fnx, vz)) b -= step;
        }
        hi = inc ? b : hi;
        lo = inc ? lo : b;
    }
    if (lo > hi) return 0;
    const int w = hi - lo + 1;
    return (w >= 64) ? ~0ull : (((1ull << w) - 1) << lo);
}

__device__ inline uint64_t raster_row(const Params &P, const ShapeDev &sh, const Pose &ps, int row) {
    PosedShape o;
    pose_shape(P, sh, ps, o);
    if (row < o.i_lo || row > o.i_hi || o.j_hi < o.j_lo) return 0;
    return raster_row_posed(P, o.n_faces, o.nx, o.nz, o.cx, o.cz, o.inv_nx, o.j_lo, o.j_hi, row);
}

}  // namespace bw

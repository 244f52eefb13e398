/*
 * bridges_b200 -- C ABI of the B200-native batched assembly_gym environment step.
 *
 * The reference (syghmon/bridges-with-reinforcement-learning) has no FFI boundary: its
 * boundary is the Python object API of
 *     assembly_gym/assembly_gym/envs/gym_env.py      (AssemblyGym, Action)
 *     assembly_gym/assembly_gym/envs/assembly_env.py (AssemblyEnv, Shape, Block)
 *     assembly_gym/assembly_gym/utils/rendering.py   (render_blocks_2d)
 *     robotoddler/utils/actions.py                   (generate_actions, filter_actions)
 * Each entry point below names the reference interface it replaces (file:line relative
 * to the reference root).  The Python adapter in bridges_b200/ binds these with ctypes
 * (see INTEGRATION.md for the stub a maintainer of the reference would add).
 *
 * Conventions
 *   - plain C, no exceptions: every call returns BW_OK (0) or a negative bw_status;
 *     bw_last_error() gives the message of the last failure on that handle.
 *   - one handle per GPU; a handle is not thread-safe, distinct handles are independent.
 *   - pointers named d_* are DEVICE pointers owned by the caller; pointers named h_* are
 *     HOST pointers.  All work is enqueued on the handle's stream; calls with d_* arguments
 *     only are asynchronous, calls with h_* arguments return after their copies completed.
 *   - lengths: E = number of lock-step environments of the handle.
 *   - all geometry is float64; observation tensors are float32 (what the Q-network eats).
 */
#ifndef BRIDGES_B200_H
#define BRIDGES_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BW_ABI_VERSION 6

/* compile-time capacities (reference configs: <= 15 blocks, <= 7 obstacles, <= 3 targets) */
#define BW_MAX_BLOCKS 16
#define BW_MAX_FACES 6      /* 2-D faces per shape (hexagon) */
#define BW_MAX_VERTS 6
#define BW_MAX_SHAPES 8
#define BW_MAX_OBSTACLES 8
#define BW_MAX_TARGETS 4
#define BW_MAX_INTERFACES 48
#define BW_IMG 64           /* observation rasters are BW_IMG x BW_IMG (successor_dqn.py:585) */

typedef enum {
    BW_OK = 0,
    BW_ERR_INVALID = -1,   /* bad argument */
    BW_ERR_CUDA = -2,      /* CUDA runtime failure (message has the CUDA error string) */
    BW_ERR_CAPACITY = -3,  /* a compile-time capacity above would be exceeded */
    BW_ERR_STATE = -4      /* call order (e.g. step before load_shapes) */
} bw_status;

typedef struct bw_handle bw_handle;

/* Constructor arguments of AssemblyEnv (assembly_env.py:164) and AssemblyGym (gym_env.py:116)
 * plus the raster window of the training script (successor_dqn.py:615-616). */
typedef struct {
    int32_t num_envs;
    int32_t device;          /* CUDA device ordinal */
    int32_t max_steps;       /* AssemblyGym(max_steps); 0 = None */
    int32_t use_caller_stream; /* 1: enqueue on `stream` below (NULL = the legacy default stream);
                                  0: the library creates its own non-blocking stream */
    double xlim[2];          /* raster / bounds window in x (default -3, 7) */
    double ylim[2];          /* raster / bounds window in z (default 0, 10) */
    double floor_halfwidth;  /* support Box half width: 0.5*(bounds[1][0]-bounds[0][0]) = 5 (assembly_env.py:290) */
    double floor_depth;      /* bounds[1][1]-bounds[0][1] = 10 */
    double mu;               /* friction coefficient, default 0.8 */
    double density;          /* default 1.0 */
    double tmax;             /* interface coplanarity tolerance (compas_cra default 1e-6) */
    double amin;             /* minimum interface area, 0.001 (assembly_env.py:304) */
    double stable_tol;       /* verdict threshold on the relative equilibrium residual, default 1e-6 */
    void *stream;            /* cudaStream_t to enqueue on when use_caller_stream = 1 */
    /* AssemblyEnv(pybullet_env=...): 0 = no physics client, every collision flag is constant False
     * (assembly_env.py:310-312, the training default); 1 = the flags of _check_collision
     * (assembly_env.py:346-391) for the last block: bounds test on the block position (:360) and
     * penetration deeper than collision_tol against blocks / floor / obstacles, computed as exact
     * convex-polygon penetration depths (separating-axis form) instead of Bullet contact points */
    int32_t collision_mode;
    int32_t reserved0;
    double collision_tol;    /* tol of _check_collision, default 0.005 */
    double bounds_lo[3];     /* AssemblyEnv.bounds[0], default (-3, -3, -1) (assembly_env.py:164-168) */
    double bounds_hi[3];     /* AssemblyEnv.bounds[1], default (7, 7, 9) */
} bw_config;

/* One entry of the block library: what Shape.from_urdf (assembly_env.py:54-68) extracts,
 * reduced to the xz-plane.  Face k is the reference's 2-D face index k (Action.face). */
typedef struct {
    int32_t n_faces;
    int32_t n_verts;
    uint32_t target_faces_mask;     /* Shape.target_faces_2d as a bit mask (bit k = face k) */
    uint32_t receiving_faces_mask;  /* Shape.receiving_faces_2d (kept for completeness; placed
                                       blocks offer all faces, assembly_env.py:153) */
    double face_nx[BW_MAX_FACES], face_nz[BW_MAX_FACES];   /* outward unit normals */
    double face_cx[BW_MAX_FACES], face_cz[BW_MAX_FACES];   /* face centres */
    double end0_x[BW_MAX_FACES], end0_z[BW_MAX_FACES];     /* face end points */
    double end1_x[BW_MAX_FACES], end1_z[BW_MAX_FACES];
    double vert_x[BW_MAX_VERTS], vert_z[BW_MAX_VERTS];     /* polygon (Shape.vertices_2d order) */
    double com_x, com_z;            /* area centroid = centre of mass of the prism */
    double area;                    /* polygon area; weight = density*area*depth */
    double depth;                   /* extent along y */
} bw_shape_desc;

/* Action dataclass, gym_env.py:102-110 */
typedef struct {
    int32_t target_block;   /* -1 = floor */
    int32_t target_face;
    int32_t shape;          /* -1 = no placement (evaluate the current assembly only) */
    int32_t face;
    double offset_x;
    double offset_y;
    int32_t frozen;         /* ignored by step (forced True, gym_env.py:238) */
    int32_t reserved0;
} bw_action;

/* A posed block: canonical pose (x, z, cos, sin) of DESIGN.md section 4 */
typedef struct {
    double x, z, c, s;
    int32_t shape;
    int32_t is_static;
} bw_block;

/* Arguments of AssemblyGym.reset (gym_env.py:255-289) for one environment */
typedef struct {
    int32_t n_obstacles, n_targets, n_blocks, reserved0;
    double obstacle_xz[BW_MAX_OBSTACLES][2];
    double target_xz[BW_MAX_TARGETS][2];
    bw_block blocks[BW_MAX_BLOCKS];     /* pre-placed blocks (reset(blocks=...)) */
} bw_task;

/* What AssemblyGym.step returns besides the images (gym_env.py:218-253) plus the pair
 * returned by stabilities_freezing (gym_env.py:325-333) */
typedef struct {
    double residual;              /* relative equilibrium residual r of the frozen solve (verdict margin) */
    double residual_unfrozen;     /* same, nothing frozen */
    double distance_to_targets[BW_MAX_TARGETS];   /* gym_env.py:154-160 (inf when no block) */
    float reward;                 /* sparse_reward, gym_env.py:11-22 */
    float lin_reward;             /* successor_dqn.py:397-401 (uses the reward image rendered by reset) */
    int32_t n_blocks;
    int32_t n_interfaces;
    int32_t newton_iters;         /* Newton steps of both solves */
    int32_t solver_kflops;        /* work estimate of both solves, in 1e3 flops */
    uint8_t stable;               /* obs['stable']: verdict with only the new block frozen */
    uint8_t stable_unfrozen;      /* stabilities_freezing()[1]: last block released */
    uint8_t collision;            /* constant 0 with collision_mode = 0 (assembly_env.py:310-312) */
    uint8_t collision_block, collision_obstacle, collision_floor, collision_boundary;
    uint8_t terminated;           /* gym_env.py:141-144 */
    uint8_t truncated;            /* max_steps reached */
    uint8_t solver_status;        /* bit0 / bit1: frozen / unfrozen solve did not converge (stable=None);
                                     bit2 / bit3: frozen / unfrozen verdict implied without (finishing) its own solve
                                     (released-block equilibrium implies frozen-block equilibrium, within a step and
                                     from the previous step's released verdict): residual is NaN or the implying one;
                                     bit4 / bit5: frozen / unfrozen verdict certified by the warm-started LP path of a
                                     real step (residual = ||b - A f|| of its basic solution when stable, NaN for a
                                     certificate of infeasibility; newton_iters does not count its pivots) */
    uint8_t error;                /* 1 = invalid action indices, 2 = capacity exceeded */
    uint8_t n_targets_reached;
    uint8_t lp_pivots;            /* simplex pivots of the LP verdict path in this step (both problems, saturating) */
    uint8_t reserved1[3];
} bw_step_out;

/* Optional observation outputs of a step; any pointer may be NULL.  The raster is
 * render_blocks_2d(obs['blocks']) (rendering.py:105-113, a bool image) either as the float32
 * tensor get_state_features makes of it (successor_dqn.py:63), as one byte per pixel, or bit-packed. */
typedef struct {
    float *block_img_f32;   /* [E,1,64,64] */
    uint8_t *block_img_u8;  /* [E,64,64], 0/1 */
    float *binary;          /* [E,6] = (stable, collision, collision_block, _obstacle, _floor, _boundary) */
    uint64_t *block_bits;   /* [E,64] the same raster bit-packed: bit x of word r = pixel (row r, column x);
                               512 B instead of 4 KB / 16 KB per environment for consumers that can take it */
} bw_obs_out;

/* One contact interface with its two contact points and their forces (bw_get_forces) */
typedef struct {
    int32_t body_a, body_b;       /* -1 = floor; a < b */
    int32_t face_a, face_b;
    double nx, nz;                /* contact normal = outward normal of a's face */
    double p0x, p0z, p1x, p1z;    /* contact points */
    double fn0, ft0, fn1, ft1;    /* min-norm equilibrium forces acting on b at p0 / p1 */
} bw_interface;

/* ---- life cycle -------------------------------------------------------------------- */
int bw_abi_version(void);
/* AssemblyEnv(...) + AssemblyGym(...) constructors */
int bw_create(const bw_config *cfg, bw_handle **out);
void bw_destroy(bw_handle *h);
const char *bw_last_error(const bw_handle *h);
void bw_config_default(bw_config *cfg);
/* cudaStreamSynchronize on the handle's stream */
int bw_sync(bw_handle *h);
/* AssemblyGym.shapes: the library the Action.shape index refers to (host pointer) */
int bw_load_shapes(bw_handle *h, const bw_shape_desc *h_shapes, int32_t n);
/* the 0.6 cube drawn for obstacles and target markers (Shape('shapes/cube06.urdf'),
 * gym_env.py:277, successor_dqn.py:73); default: an exact 0.6 x 0.6 axis-aligned box */
int bw_set_marker_shape(bw_handle *h, const bw_shape_desc *h_shape);
/* normalised 1-D Gaussian (101 float32 taps) of get_task_features / convolve_with_gaussian
 * (successor_dqn.py:78-80, robotoddler/utils/utils.py:93-114); default: sigma 16 */
int bw_set_task_kernel(bw_handle *h, const float *h_kernel1d, int32_t n);
/* per-environment friction coefficient (AssemblyEnv.mu); host array of E doubles */
int bw_set_mu(bw_handle *h, const double *h_mu);

/* ---- reset: AssemblyGym.reset, gym_env.py:255-289 ---------------------------------------
 * d_tasks: E tasks, or NULL to keep every env's obstacles/targets and only clear its blocks.
 * d_mask:  E bytes, env e is reset iff d_mask[e] != 0; NULL = all.
 * Also renders the obstacle raster and the Gaussian-blurred target raster
 * (get_task_features, successor_dqn.py:67-85). */
int bw_reset(bw_handle *h, const bw_task *d_tasks, const uint8_t *d_mask);
int bw_reset_host(bw_handle *h, const bw_task *h_tasks, const uint8_t *h_mask);
/* reset exactly the envs whose last step returned terminated|truncated (keeps their task) */
int bw_reset_done(bw_handle *h);

/* ---- step: AssemblyGym.step (gym_env.py:218-253) + stabilities_freezing (:325-333) ------
 * d_actions[E], d_mask (NULL = all), d_out[E]; `obs` (may be NULL) holds DEVICE pointers for
 * bw_step and HOST pointers for bw_step_host (get_state_features, successor_dqn.py:47-64). */
int bw_step(bw_handle *h, const bw_action *d_actions, const uint8_t *d_mask, bw_step_out *d_out,
            const bw_obs_out *obs);
int bw_step_host(bw_handle *h, const bw_action *h_actions, const uint8_t *h_mask, bw_step_out *h_out,
                 const bw_obs_out *obs);
/* The same without an action: interfaces, both verdicts (stabilities_freezing, gym_env.py:325-333), distances and
 * observations of the assemblies as they stand -- AssemblyEnv._update_state_info (assembly_env.py:404-438) for every
 * environment, what bw_step does for Action.shape = -1, from a smaller kernel image (no placement, no raster update,
 * no LP path).  d_mask (NULL = all), d_out[E], obs as in bw_step (device pointers, may be NULL). */
int bw_evaluate(bw_handle *h, const uint8_t *d_mask, bw_step_out *d_out, const bw_obs_out *obs);
/* How bw_step_host moves its buffers.  0 (default): automatic -- when every host buffer of the call
 * is pinned (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory()) the step kernel reads the
 * actions from and writes records / images to host memory directly over PCIe, every environment
 * as soon as it is done, so the transfers overlap the solves; pageable buffers are staged through
 * device buffers with cudaMemcpyAsync.  1: always stage. */
int bw_set_host_transfer(bw_handle *h, int32_t mode);

/* ---- observations: _get_obs + get_state_features / get_task_features ----------------
 * any pointer may be NULL.  Images are [E,1,64,64] f32, row 0 = top (rendering.py:105-113). */
int bw_observe(bw_handle *h, float *d_block_img, float *d_binary, float *d_obstacle_img, float *d_reward_img);
int bw_observe_host(bw_handle *h, float *h_block_img, float *h_binary, float *h_obstacle_img, float *h_reward_img);

/* ---- candidate actions: generate_actions + get_action_features + filter_actions --------
 * (robotoddler/utils/actions.py:7-82, successor_dqn.py:88-94).  Candidates are written in
 * the reference's generation order, Amax per env; d_n_cand[e] of them are meaningful.
 * d_valid[e,a] = 1 iff the candidate survives filter_actions (bounds test of
 * collision_on_action gym_env.py:304-323, no raster overlap with blocks / obstacles).
 * d_action_bits (optional): [E,Amax,64] u64, bit x of word r = pixel (row r, col x).
 * The handle keeps the placement, bounds flag and raster rows of every possible candidate between
 * calls (device memory: E x groups x (n_ground + max_blocks*6*n_offsets) x 516 B, allocated by the first
 * call; results are identical to recomputing them).  The environment variable BW_CAND_CACHE_MB, read by
 * bw_create, bounds that memory (default 4096; 0 = always recompute). */
int bw_enumerate_actions(bw_handle *h, const double *h_x_discr_ground, int32_t n_ground,
                         const double *h_offset_values, int32_t n_offsets, int32_t amax,
                         bw_action *d_cand, uint8_t *d_valid, int32_t *d_n_cand, uint64_t *d_action_bits);
/* The same enumeration WITHOUT the raster copies (filter_actions keeps a fraction of the candidates, a policy looks
 * at those): d_slot[e,a] (i32 [E,Amax]) = where the raster of candidate a lives in the handle's candidate store, -1
 * = none (never for a valid candidate).  bw_gather_action_bits copies the rasters of chosen candidates out:
 * d_bits[i] ([n,64] u64) = raster of candidate d_index[i] of environment d_env[i] (d_env NULL: environment i).
 * Slots are good until the next enumeration or state change of their environment.  Candidates that were listed by
 * the previous call are only tested against the pixels the block raster gained since then (results identical to
 * bw_enumerate_actions).  BW_ERR_CAPACITY when the store is switched off (BW_CAND_CACHE_MB = 0 / no memory). */
int bw_enumerate_actions_stored(bw_handle *h, const double *h_x_discr_ground, int32_t n_ground,
                                const double *h_offset_values, int32_t n_offsets, int32_t amax,
                                bw_action *d_cand, uint8_t *d_valid, int32_t *d_n_cand, int32_t *d_slot);
int bw_gather_action_bits(bw_handle *h, const int32_t *d_slot, int32_t amax, const int32_t *d_env,
                          const int32_t *d_index, int64_t n, uint64_t *d_bits);
/* generate_actions is unbounded, the buffers above hold amax candidates per environment: lists that did not
 * fit are cut to amax (d_n_cand[e] = amax) and remembered.  *h_needed = the largest untruncated count any
 * environment had since the last query (0 = every list was complete); synchronises and resets the mark. */
int bw_candidate_overflow(bw_handle *h, int32_t *h_needed);
/* expand bit rasters [n,64] u64 -> [n,1,64,64] f32 */
int bw_expand_bits(bw_handle *h, const uint64_t *d_bits, int64_t n, float *d_img);
/* synthetic policy for benchmarks/tests: pick for every env a uniformly random valid candidate
 * (counter-based hash of seed, env, step); envs without a valid candidate get shape = -1 and are flagged as
 * done (rollout_episode ends the episode when no candidate is left, successor_dqn.py:409-411), so that the
 * next bw_reset_done starts them afresh. */
int bw_select_random(bw_handle *h, const bw_action *d_cand, const uint8_t *d_valid, const int32_t *d_n_cand,
                     int32_t amax, uint64_t seed, bw_action *d_actions, int32_t *d_index);

/* ---- fused lock-step rollout: rollout_episode (successor_dqn.py:365-475) for E environments at once, with the
 * transitions (Transition, successor_dqn.py:27-44) written as packed records into a caller-owned device ring that
 * plays the part of ReplayBuffer.memory (robotoddler/utils/replay_memory.py:10-43).  One iteration =
 *     policy picks one valid candidate per environment  ->  step + stabilities_freezing + lin_reward  ->
 *     record  ->  auto-reset of finished episodes  ->  candidates of the next states (generate_actions +
 *     get_action_features + filter_actions)  ->  done |= "no candidate left" (successor_dqn.py:409-411)
 * as a fixed sequence of kernels on the handle's stream: no host synchronisation, no per-step host work besides
 * the launches.  Rasters stay bit-packed (bit x of word r = pixel (row r, column x)). */
typedef struct {
    uint64_t block_bits[BW_IMG];       /* block_features: render_blocks_2d of the state before the step */
    uint64_t action_bits[BW_IMG];      /* action_features: raster of the chosen candidate */
    uint64_t next_block_bits[BW_IMG];  /* next_block_features */
    bw_action action;                  /* the chosen Action */
    float reward;                      /* sparse_reward (gym_env.py:11-22) */
    float lin_reward;                  /* successor_dqn.py:397-401 */
    int32_t env;                       /* global environment index (env_id_base + e) */
    int32_t step;                      /* rollout iteration counter of the handle */
    int32_t n_next_candidates;         /* valid candidates of the next state (meaningful when the episode goes on) */
    uint8_t binary;                    /* binary_features of the state before, bit k = feature k (stable, collision,
                                          collision_block, _obstacle, _floor, _boundary) */
    uint8_t next_binary;
    uint8_t done;                      /* terminated | truncated | no candidate left in the next state */
    uint8_t terminated, truncated;
    uint8_t stable, stable_unfrozen;   /* stabilities_freezing() of the new state */
    uint8_t valid;                     /* 0: this environment had no candidate in this iteration: not a transition */
    uint8_t reserved[4];
} bw_transition;                       /* 1608 bytes */

/* device views of the candidate buffers the handle keeps for the rollout (valid until the next rollout call) */
typedef struct {
    const bw_action *cand;       /* [E,amax] in generate_actions order */
    const uint8_t *valid;        /* [E,amax] filter_actions mask */
    const int32_t *n_cand;       /* [E] */
    const int32_t *n_valid;      /* [E] */
    const uint64_t *action_bits; /* [E,amax,64] dense rasters -- NULL when the candidates live in the handle's store */
    const int32_t *slot;         /* [E,amax] store slots (NULL without a store); read rasters with bw_rollout_gather_bits */
    int32_t amax;
    int32_t reserved0;
} bw_rollout_view;

/* arguments of generate_actions for the rollout; allocates the handle-owned candidate buffers.
 * env_id_base: added to the environment index in the records (rank * E on a sharded run). */
int bw_rollout_configure(bw_handle *h, const double *h_x_discr_ground, int32_t n_ground, const double *h_offset_values,
                         int32_t n_offsets, int32_t amax, int32_t env_id_base);
/* candidates of the CURRENT states (first call, or after bw_reset / bw_step changed them behind the rollout's
 * back; bw_rollout_commit leaves the candidates of the next states behind by itself).  out may be NULL. */
int bw_rollout_begin(bw_handle *h, bw_rollout_view *out);
/* one iteration with the caller's policy: d_index[e] = index of the chosen candidate of environment e (ignored
 * where n_valid[e] = 0).  d_slots: E records (record e belongs to environment e).  obs: optional observation
 * outputs of the step as in bw_step (device pointers), may be NULL. */
int bw_rollout_commit(bw_handle *h, const int32_t *d_index, bw_transition *d_slots, const bw_obs_out *obs);
/* rasters of chosen candidates of the CURRENT candidate lists: d_bits[i] ([n,64] u64) = raster of candidate d_index[i]
 * of environment d_env[i] (d_env NULL: environment i); works with and without a candidate store */
int bw_rollout_gather_bits(bw_handle *h, const int32_t *d_env, const int32_t *d_index, int64_t n, uint64_t *d_bits);
/* n_steps iterations with the built-in uniformly random policy (the synthetic policy of the benchmarks).  The
 * records of iteration k go to d_ring[(start + k*E + e) % capacity]; capacity must be a multiple of E. */
int bw_rollout_random(bw_handle *h, int32_t n_steps, uint64_t seed, bw_transition *d_ring, int64_t capacity,
                      int64_t start);
/* ReplayBuffer.sample(stack_tensors=True) for packed records: expands the records d_ring[d_indices[i]] into the
 * learner's tensors; any output may be NULL.  Images [n,1,64,64] f32, binary [n,6] f32, reward / lin_reward [n]
 * f32, done [n] u8. */
int bw_unpack_transitions(bw_handle *h, const bw_transition *d_ring, const int64_t *d_indices, int64_t n,
                          float *d_block, float *d_action, float *d_next_block, float *d_binary, float *d_next_binary,
                          float *d_reward, float *d_lin_reward, uint8_t *d_done);

/* ---- state read-back (adapters, checkpointing, tests) ------------------------------ */
int bw_get_state(bw_handle *h, bw_block *h_blocks /*[E,BW_MAX_BLOCKS]*/, int32_t *h_n_blocks /*[E]*/);
int bw_get_raster_bits(bw_handle *h, uint64_t *h_block_bits /*[E,64]*/, uint64_t *h_obstacle_bits /*[E,64]*/);
/* the same into caller-owned DEVICE buffers (device-to-device, asynchronous on the handle's stream) */
int bw_copy_raster_bits(bw_handle *h, uint64_t *d_block_bits /*[E,64]*/, uint64_t *d_obstacle_bits /*[E,64]*/);
/* targets_remaining / targets_reached of AssemblyGym (gym_env.py:162-168) as indices into the
 * task's target list, in list order; h_counts[e] = {n_remaining, n_reached} */
int bw_get_target_state(bw_handle *h, int8_t *h_remaining /*[E,BW_MAX_TARGETS]*/,
                        int8_t *h_reached /*[E,BW_MAX_TARGETS]*/, int32_t *h_counts /*[E,2]*/);
/* AssemblyGym.create_block (gym_env.py:204-216) + collision_on_action (gym_env.py:304-323)
 * for one hypothetical action per env, without touching the state.  h_blocks[e] = posed block;
 * h_flags[e]: bit0 invalid indices, bit1 env full, bit2 a vertex leaves xlim/ylim (+-1e-6) or
 * dips below z = -1e-6 */
int bw_query_placement_host(bw_handle *h, const bw_action *h_actions, const double *xlim2, const double *ylim2,
                            bw_block *h_blocks, uint8_t *h_flags);
/* render_blocks_2d (rendering.py:105-113) for an arbitrary list of posed blocks, 64 x 64 only:
 * blocks[i].shape indexes h_shapes; h_bits[64], bit x of word r = pixel (row r, col x) */
int bw_render_blocks_host(bw_handle *h, const bw_shape_desc *h_shapes, int32_t n_shapes, const bw_block *h_blocks,
                          int32_t n_blocks, const double *xlim2, const double *ylim2, uint64_t *h_bits);
/* Shape.contains_2d / Block.contains_2d (assembly_env.py:126-137) for n arbitrary points (x, z): h_inside[i] = 1 iff
 * the point lies in every half-plane of the shape posed at h_block (NULL = the unposed shape of the library file).
 * The drop-in render_blocks_2d uses it for image sizes other than 64 x 64 (rendering.py:105-113). */
int bw_contains_2d_host(bw_handle *h, const bw_shape_desc *h_shape, const bw_block *h_block, const double *h_points_xz,
                        int64_t n, uint8_t *h_inside);
/* interfaces and min-norm contact forces of the last step / evaluation (frozen variant) */
/* variant 0: supports as in the last step's verdict (new block frozen); 1: last block released */
int bw_get_forces(bw_handle *h, int32_t variant, bw_interface *h_itf /*[E,BW_MAX_INTERFACES]*/, int32_t *h_n_itf /*[E]*/);
/* AssemblyEnv.freeze_block / unfreeze_block (assembly_env.py:404-438): host array [E] of
 * bit masks (bit i = block i is a support) used by the next bw_step(shape=-1) evaluation */
int bw_set_static_mask(bw_handle *h, const uint32_t *h_mask);

/* ---- measurement helpers ------------------------------------------------------------- */
/* bw_set_timing(h, 1) brackets the step kernel of bw_step with CUDA events on the handle's
 * stream; bw_last_step_kernel_ms then returns (after synchronising) in h_ms2[0] the elapsed ms
 * of that kernel (placement, interfaces, two solves, bookkeeping, raster update, f32
 * observation write); h_ms2[1] is reserved (0) */
int bw_set_timing(bw_handle *h, int32_t enabled);
int bw_last_step_kernel_ms(bw_handle *h, float *h_ms2);
/* sustained FP64 FMA throughput of the device (GFLOP/s), measured by a micro-benchmark:
 * the denominator of the solver's compute roofline */
int bw_fp64_peak_gflops(bw_handle *h, double *h_gflops);
/* counters of the LP verdict path (collected only when the handle was created with BW_LP_STATS set in the
 * environment -- a tuning hook, not part of the reference interface): h_stats[32], layout in csrc/bw_step.cu */
int bw_debug_lp_stats(bw_handle *h, uint64_t *h_stats);
/* number of kernels launched by this handle so far */
int64_t bw_kernel_launches(const bw_handle *h);

#ifdef __cplusplus
}
#endif
#endif /* BRIDGES_B200_H */

"""Lock-step batch of assembly environments on one GPU.

`BatchedAssemblyGym` is the batched form of `AssemblyGym` (gym_env.py:112-333): E
independent assemblies advance with one `step` call.  It owns one `bw_handle` (one CUDA
stream) and hands tensors over as torch CUDA tensors; torch is used for device memory
only.  The single-environment drop-in classes in `envs/gym_env.py` are thin views of it.
"""
import ctypes as C

import numpy as np
import torch

from .. import lib as L
from ..shapes_io import ShapeTables, load_shape_tables


def shape_desc(tables, target_faces=None, receiving_faces=None):
    """ShapeTables -> bw_shape_desc."""
    d = L.bw_shape_desc()
    nf, nv = len(tables.normals), len(tables.polygon)
    if nf > L.BW_MAX_FACES or nv > L.BW_MAX_VERTS:
        raise L.BridgesError("shape exceeds BW_MAX_FACES / BW_MAX_VERTS")
    d.n_faces, d.n_verts = nf, nv
    full = (1 << nf) - 1
    d.target_faces_mask = full if not target_faces else sum(1 << int(f) for f in target_faces)
    d.receiving_faces_mask = full if not receiving_faces else sum(1 << int(f) for f in receiving_faces)
    for k in range(nf):
        d.face_nx[k], d.face_nz[k] = tables.normals[k]
        d.face_cx[k], d.face_cz[k] = tables.centers[k]
        (d.end0_x[k], d.end0_z[k]), (d.end1_x[k], d.end1_z[k]) = tables.ends[k]
    for k in range(nv):
        d.vert_x[k], d.vert_z[k] = tables.polygon[k]
    d.com_x, d.com_z = tables.com
    d.area, d.depth = tables.area, tables.depth
    return d


def gaussian_kernel1d(kernel_size=101, sigma=16):
    """robotoddler/utils/utils.py:93-101, the 1-D factor (float32, torch arithmetic)."""
    coords = torch.arange(kernel_size) - kernel_size // 2
    k = torch.exp(-(coords.float() ** 2) / (2 * sigma ** 2))
    k /= k.sum()
    return k


class CandidateRasters:
    """Bit rasters of listed candidates that live in the handle's candidate store (or in its dense copies):
    `r[env_idx, cand_idx]` (two equally long integer CUDA tensors) -> int64 [n, 64], `r.dense()` -> [E, amax, 64].
    Good until the next enumeration / rollout iteration."""

    def __init__(self, env, E, amax, gather, dense=None):
        self.env, self.E, self.amax, self._gather, self._dense = env, E, amax, gather, dense

    def __getitem__(self, key):
        er, ar = key
        if self._dense is not None:        # no candidate store: the handle copied the rasters out
            dev = self.env.device
            return self._dense[torch.as_tensor(er, device=dev).long(), torch.as_tensor(ar, device=dev).long()]
        er = torch.as_tensor(er, device=self.env.device).to(torch.int32).contiguous()
        ar = torch.as_tensor(ar, device=self.env.device).to(torch.int32).contiguous()
        if er.shape != ar.shape or er.dim() != 1:
            raise L.BridgesError("CandidateRasters[env_idx, cand_idx]: two 1-D index tensors of equal length")
        out = torch.empty((er.numel(), L.BW_IMG), dtype=torch.int64, device=self.env.device)
        self.env._check(self._gather(er.data_ptr(), ar.data_ptr(), er.numel(), out.data_ptr()))
        return out

    def dense(self):
        if self._dense is not None:
            return self._dense
        dev = self.env.device
        er = torch.arange(self.E, device=dev, dtype=torch.int32).repeat_interleave(self.amax)
        ar = torch.arange(self.amax, device=dev, dtype=torch.int32).repeat(self.E)
        return self[er, ar].reshape(self.E, self.amax, L.BW_IMG)


class BatchedAssemblyGym:
    def __init__(self, num_envs, shapes, max_steps=None, device=0, mu=0.8, density=1.0, xlim=(-3.0, 7.0),
                 ylim=(0.0, 10.0), bounds=None, tmax=1e-6, amin=1e-3, stable_tol=1e-6, stream=None,
                 collision=False, collision_tol=0.005):
        """collision=True: the flags of `AssemblyEnv._check_collision` (assembly_env.py:346-391, what the
        reference produces with pybullet_env=True) from exact polygon penetration depths; False: constant
        False flags (the reference without a physics client, assembly_env.py:310-312)."""
        if not torch.cuda.is_available():
            raise L.BridgesError("bridges_b200 needs a CUDA device (no CPU fallback)")
        self.lib = L.load()
        self.dt = L.np_dtypes()
        self.num_envs = int(num_envs)
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        if bounds is None:
            bounds = np.array([[-3.0, -3.0, -1.0], [7.0, 7.0, 9.0]])
        self.bounds = np.asarray(bounds, dtype=float)
        self.xlim, self.ylim = tuple(map(float, xlim)), tuple(map(float, ylim))
        self.max_steps = max_steps
        cfg = L.bw_config()
        self.lib.bw_config_default(C.byref(cfg))
        cfg.num_envs = self.num_envs
        cfg.device = self.device.index
        cfg.max_steps = int(max_steps) if max_steps else 0
        cfg.xlim[0], cfg.xlim[1] = self.xlim
        cfg.ylim[0], cfg.ylim[1] = self.ylim
        cfg.floor_halfwidth = 0.5 * float(self.bounds[1][0] - self.bounds[0][0])
        cfg.floor_depth = float(self.bounds[1][1] - self.bounds[0][1])
        cfg.mu, cfg.density, cfg.tmax, cfg.amin, cfg.stable_tol = mu, density, tmax, amin, stable_tol
        cfg.collision_mode = 1 if collision else 0
        cfg.collision_tol = float(collision_tol)
        for k in range(3):
            cfg.bounds_lo[k], cfg.bounds_hi[k] = float(self.bounds[0][k]), float(self.bounds[1][k])
        # enqueue on torch's current stream unless told otherwise, so that tensor fills / copies
        # issued through torch and the library's kernels are ordered without extra syncs
        with torch.cuda.device(self.device):
            cfg.stream = stream if stream is not None else torch.cuda.current_stream().cuda_stream
            cfg.use_caller_stream = 1          # 0/NULL is torch's (legacy default) stream, a valid choice
        self.handle = C.c_void_p()
        rc = self.lib.bw_create(C.byref(cfg), C.byref(self.handle))
        self._check(rc)
        self.mu, self.density = mu, density
        self.set_shapes(shapes)
        marker = shape_desc(load_shape_tables("shapes/cube06.urdf"))
        self._check(self.lib.bw_set_marker_shape(self.handle, C.byref(marker)))
        k = gaussian_kernel1d().numpy().astype(np.float32)
        self._check(self.lib.bw_set_task_kernel(self.handle, k.ctypes.data, k.size))
        E = self.num_envs
        self._actions = torch.zeros(E * self.dt["action"].itemsize, dtype=torch.uint8, device=self.device)
        self._out = torch.zeros(E * self.dt["step_out"].itemsize, dtype=torch.uint8, device=self.device)
        self._cand = None

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc):
        L.check(self.lib, self.handle, rc)

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.bw_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self._check(self.lib.bw_sync(self.handle))

    def _to_device_bytes(self, array):
        host = torch.from_numpy(np.ascontiguousarray(array).view(np.uint8).reshape(-1))
        return host.to(self.device)

    def set_shapes(self, shapes):
        """`shapes`: objects with `.tables` (envs.assembly_env.Shape), ShapeTables or urdf paths."""
        descs = (L.bw_shape_desc * len(shapes))()
        self.shape_tables, self._target_masks = [], []
        for i, s in enumerate(shapes):
            if isinstance(s, str):
                tables, tf, rf = load_shape_tables(s), None, None
            elif isinstance(s, ShapeTables):
                tables, tf, rf = s, None, None
            else:
                tables, tf, rf = s.tables, s._target_faces_2d, s._receiving_faces_2d
            self.shape_tables.append(tables)
            descs[i] = shape_desc(tables, tf, rf)
            self._target_masks.append(int(descs[i].target_faces_mask))
        self._check(self.lib.bw_load_shapes(self.handle, descs, len(shapes)))

    def set_mu(self, mu):
        mu = np.ascontiguousarray(np.broadcast_to(np.asarray(mu, dtype=np.float64), (self.num_envs,)))
        self._check(self.lib.bw_set_mu(self.handle, mu.ctypes.data))

    def set_static_mask(self, mask):
        mask = np.ascontiguousarray(np.broadcast_to(np.asarray(mask, dtype=np.uint32), (self.num_envs,)))
        self._check(self.lib.bw_set_static_mask(self.handle, mask.ctypes.data))

    # ------------------------------------------------------------------ reset
    def make_tasks(self, tasks):
        """list of dict(obstacles=[(x,y,z)], targets=[(x,y,z)], blocks=[(x,z,c,s,shape)]) or one dict."""
        if isinstance(tasks, dict):
            tasks = [tasks] * self.num_envs
        arr = np.zeros(self.num_envs, dtype=self.dt["task"])
        for e, t in enumerate(tasks):
            obstacles, targets, blocks = t.get("obstacles") or [], t.get("targets") or [], t.get("blocks") or []
            if len(obstacles) > L.BW_MAX_OBSTACLES or len(targets) > L.BW_MAX_TARGETS or len(blocks) > L.BW_MAX_BLOCKS:
                raise L.BridgesError("task exceeds BW_MAX_OBSTACLES / BW_MAX_TARGETS / BW_MAX_BLOCKS")
            arr[e]["n_obstacles"], arr[e]["n_targets"], arr[e]["n_blocks"] = len(obstacles), len(targets), len(blocks)
            for i, p in enumerate(obstacles):
                arr[e]["obstacle_xz"][i] = (p[0], p[2])
            for i, p in enumerate(targets):
                arr[e]["target_xz"][i] = (p[0], p[2])
            for i, b in enumerate(blocks):
                arr[e]["blocks"][i] = (b[0], b[1], b[2], b[3], int(b[4]), 0)
        return arr

    def reset(self, tasks=None, mask=None):
        """mask: host array or a uint8 CUDA tensor [E] (then, without tasks, nothing synchronises)."""
        d_tasks = self._to_device_bytes(self.make_tasks(tasks)) if tasks is not None else None
        if isinstance(mask, torch.Tensor) and mask.is_cuda:
            d_mask = mask.to(torch.uint8).contiguous()
            self._check(self.lib.bw_reset(self.handle, d_tasks.data_ptr() if d_tasks is not None else None,
                                          d_mask.data_ptr()))
            if d_tasks is not None:
                self.sync()
            else:
                self._keep = d_mask                 # stays alive until the next call; same stream as torch
            return
        d_mask = self._to_device_bytes(np.asarray(mask, dtype=np.uint8)) if mask is not None else None
        self._check(self.lib.bw_reset(self.handle, d_tasks.data_ptr() if d_tasks is not None else None,
                                      d_mask.data_ptr() if d_mask is not None else None))
        self.sync()     # the staging tensors above must outlive the kernel

    def reset_done(self):
        self._check(self.lib.bw_reset_done(self.handle))

    # ------------------------------------------------------------------ step
    def actions_array(self, actions):
        """list of Action-like objects (or tuples tb, tf, shape, face, ox, oy) -> structured array."""
        arr = np.zeros(self.num_envs, dtype=self.dt["action"])
        for e, a in enumerate(actions):
            if a is None:
                arr[e] = (-1, 0, -1, 0, 0.0, 0.0, 0, 0)
            elif isinstance(a, (tuple, list)):
                t = tuple(a) + (0.0,) * (6 - len(a))
                arr[e] = (t[0], t[1], t[2], t[3], t[4], t[5], 0, 0)
            else:
                arr[e] = (a.target_block, a.target_face, a.shape, a.face, a.offset_x, a.offset_y, int(a.frozen), 0)
        return arr

    def step(self, actions, mask=None, block_img=None, binary=None, block_u8=None, block_bits=None):
        """actions: structured array / list (host) or a uint8 CUDA tensor holding bw_action[E].
        Optional outputs (CUDA tensors): block_img f32 [E,1,64,64], block_u8 [E,64,64], binary f32 [E,6],
        block_bits int64 [E,64] (bit-packed raster, see `bits_to_bool`).
        Returns the device uint8 tensor holding bw_step_out[E] (see `read_out`)."""
        if isinstance(actions, torch.Tensor):
            d_act = actions
        else:
            if not isinstance(actions, np.ndarray):
                actions = self.actions_array(actions)
            self._actions.copy_(torch.from_numpy(actions.view(np.uint8).reshape(-1)), non_blocking=False)
            d_act = self._actions
        mask_on_device = isinstance(mask, torch.Tensor) and mask.is_cuda
        if mask_on_device:
            d_mask = mask.to(torch.uint8).contiguous()
            self._keep = d_mask                      # same stream as torch: alive until the next call is enough
        else:
            d_mask = self._to_device_bytes(np.asarray(mask, dtype=np.uint8)) if mask is not None else None
        obs = None
        if block_img is not None or binary is not None or block_u8 is not None or block_bits is not None:
            obs = L.bw_obs_out(block_img.data_ptr() if block_img is not None else None,
                               block_u8.data_ptr() if block_u8 is not None else None,
                               binary.data_ptr() if binary is not None else None,
                               block_bits.data_ptr() if block_bits is not None else None)
        self._check(self.lib.bw_step(self.handle, d_act.data_ptr(), d_mask.data_ptr() if d_mask is not None else None,
                                     self._out.data_ptr(), C.byref(obs) if obs is not None else None))
        if d_mask is not None and not mask_on_device:
            self.sync()
        return self._out

    def read_out(self, out=None):
        self.sync()
        out = self._out if out is None else out
        return out.cpu().numpy().view(self.dt["step_out"])

    def out_fields(self, out=None, names=("reward", "lin_reward", "terminated", "truncated", "stable",
                                          "stable_unfrozen", "n_blocks")):
        """Fields of the bw_step_out records as CUDA tensors [E] (no host synchronisation)."""
        out = self._out if out is None else out
        rec = out.view(self.num_envs, self.dt["step_out"].itemsize)
        tmap = {"<f4": torch.float32, "<f8": torch.float64, "<i4": torch.int32, "|u1": torch.uint8}
        res = {}
        for name in names:
            fdt, off = self.dt["step_out"].fields[name][:2]
            if fdt.shape:
                raise L.BridgesError("array fields are not exposed as tensors")
            col = rec[:, off:off + fdt.itemsize].contiguous()
            res[name] = col.view(tmap[fdt.str]).reshape(self.num_envs)
        return res

    def evaluate(self, mask=None, block_img=None, binary=None, block_bits=None):
        """Verdicts, distances and (optionally) observations of the current assemblies without placing a block:
        `bw_evaluate`, what `step` does for Action.shape = -1 from the evaluation-only kernel image.
        mask: uint8 CUDA tensor [E] or None.  Returns the device tensor holding bw_step_out[E]."""
        d_mask = None
        if mask is not None:
            d_mask = mask.to(torch.uint8).contiguous() if isinstance(mask, torch.Tensor) and mask.is_cuda \
                else self._to_device_bytes(np.asarray(mask, dtype=np.uint8))
            self._keep = d_mask
        obs = None
        if block_img is not None or binary is not None or block_bits is not None:
            obs = L.bw_obs_out(block_img.data_ptr() if block_img is not None else None, None,
                               binary.data_ptr() if binary is not None else None,
                               block_bits.data_ptr() if block_bits is not None else None)
        self._check(self.lib.bw_evaluate(self.handle, d_mask.data_ptr() if d_mask is not None else None,
                                         self._out.data_ptr(), C.byref(obs) if obs is not None else None))
        if d_mask is not None and not (isinstance(mask, torch.Tensor) and mask.is_cuda):
            self.sync()
        return self._out

    # ------------------------------------------------------------------ observations
    def observe(self, block=True, binary=True, obstacle=False, reward=False):
        E = self.num_envs
        mk = lambda: torch.empty((E, 1, L.BW_IMG, L.BW_IMG), dtype=torch.float32, device=self.device)
        out = dict(block=mk() if block else None,
                   binary=torch.empty((E, 6), dtype=torch.float32, device=self.device) if binary else None,
                   obstacle=mk() if obstacle else None, reward=mk() if reward else None)
        ptr = lambda t: t.data_ptr() if t is not None else None
        self._check(self.lib.bw_observe(self.handle, ptr(out["block"]), ptr(out["binary"]), ptr(out["obstacle"]),
                                        ptr(out["reward"])))
        self.sync()
        return out

    def raster_bits(self):
        E = self.num_envs
        blk = np.zeros((E, L.BW_IMG), dtype=np.uint64)
        obs = np.zeros((E, L.BW_IMG), dtype=np.uint64)
        self._check(self.lib.bw_get_raster_bits(self.handle, blk.ctypes.data, obs.ctypes.data))
        return blk, obs

    def raster_bits_device(self, obstacles=False):
        """Bit rasters as an int64 CUDA tensor [E, 64] (device-to-device copy on the handle's stream)."""
        out = torch.empty((self.num_envs, L.BW_IMG), dtype=torch.int64, device=self.device)
        if obstacles:
            self._check(self.lib.bw_copy_raster_bits(self.handle, None, out.data_ptr()))
        else:
            self._check(self.lib.bw_copy_raster_bits(self.handle, out.data_ptr(), None))
        return out

    @staticmethod
    def bits_to_bool(bits):
        """[..., 64] uint64 rows -> [..., 64, 64] bool (row 0 = top, bit x = column x)."""
        bits = np.asarray(bits, dtype=np.uint64)
        cols = np.arange(L.BW_IMG, dtype=np.uint64)
        return ((bits[..., None] >> cols) & np.uint64(1)).astype(bool)

    # ------------------------------------------------------------------ candidate actions
    def enumerate_actions(self, x_discr_ground, offset_values=(0.0,), amax=256, with_bits=True):
        """generate_actions + get_action_features + filter_actions of every environment (actions.py:7-82,
        successor_dqn.py:88-94).  with_bits=True: dense raster copies c["bits"] int64 [E,amax,64];
        with_bits="stored": the rasters stay in the handle's candidate store, c["bits"] is a `CandidateRasters`
        (`c["bits"][env_idx, cand_idx]` gathers the ones a caller looks at, `.dense()` all of them);
        with_bits=False: no rasters."""
        E = self.num_envs
        if self._cand is None or self._cand["amax"] != amax:
            self._cand = dict(
                amax=amax,
                cand=torch.zeros(E * amax * self.dt["action"].itemsize, dtype=torch.uint8, device=self.device),
                valid=torch.zeros((E, amax), dtype=torch.uint8, device=self.device),
                n=torch.zeros(E, dtype=torch.int32, device=self.device), bits=None, dense=None, slot=None)
        c = self._cand
        g = np.ascontiguousarray(np.asarray(x_discr_ground, dtype=np.float64))
        o = np.ascontiguousarray(np.asarray(offset_values, dtype=np.float64))
        if with_bits == "stored":
            if c["slot"] is None:
                c["slot"] = torch.zeros((E, amax), dtype=torch.int32, device=self.device)
            rc = self.lib.bw_enumerate_actions_stored(
                self.handle, g.ctypes.data, g.size, o.ctypes.data, o.size, amax, c["cand"].data_ptr(),
                c["valid"].data_ptr(), c["n"].data_ptr(), c["slot"].data_ptr())
            if rc != L.BW_ERR_CAPACITY:
                self._check(rc)
                slot = c["slot"]
                c["bits"] = CandidateRasters(self, E, amax, lambda env, idx, n, out: self.lib.bw_gather_action_bits(
                    self.handle, slot.data_ptr(), amax, env, idx, n, out))
                return c
            # the handle has no candidate store (BW_CAND_CACHE_MB, device memory): dense copies behind the same interface
            dense = self.enumerate_actions(x_discr_ground, offset_values, amax=amax, with_bits=True)["bits"]
            c["bits"] = CandidateRasters(self, E, amax, None, dense=dense)
            return c
        if with_bits and c["dense"] is None:
            c["dense"] = torch.zeros((E, amax, L.BW_IMG), dtype=torch.int64, device=self.device)
        self._check(self.lib.bw_enumerate_actions(
            self.handle, g.ctypes.data, g.size, o.ctypes.data, o.size, amax, c["cand"].data_ptr(),
            c["valid"].data_ptr(), c["n"].data_ptr(), c["dense"].data_ptr() if with_bits else None))
        c["bits"] = c["dense"] if with_bits else None
        return c

    def candidate_overflow(self):
        """Largest candidate count an `enumerate_actions` call since the last query had to cut to its `amax`
        (0: every list was complete).  Synchronises."""
        need = C.c_int32(0)
        self._check(self.lib.bw_candidate_overflow(self.handle, C.byref(need)))
        return int(need.value)

    def max_candidates(self, n_ground, n_offsets=1):
        """Upper bound of the candidate count of one environment (generate_actions, actions.py:7-52): every
        (shape, target face) group offers the ground offsets plus every face of every placed block."""
        groups = sum(bin(self._target_masks[i]).count("1") for i in range(len(self.shape_tables)))
        faces = max(len(t.normals) for t in self.shape_tables)
        cap = int(self.max_steps) if self.max_steps else L.BW_MAX_BLOCKS
        return groups * (n_ground + max(cap - 1, 0) * faces * n_offsets)

    def select_random(self, seed, cand=None):
        c = cand or self._cand
        idx = torch.empty(self.num_envs, dtype=torch.int32, device=self.device)
        self._check(self.lib.bw_select_random(self.handle, c["cand"].data_ptr(), c["valid"].data_ptr(),
                                              c["n"].data_ptr(), c["amax"], int(seed) & (2 ** 64 - 1),
                                              self._actions.data_ptr(), idx.data_ptr()))
        return self._actions, idx

    def expand_bits(self, bits):
        """int64/uint64 CUDA tensor [n, 64] -> float32 [n, 1, 64, 64]."""
        n = bits.numel() // L.BW_IMG
        img = torch.empty((n, 1, L.BW_IMG, L.BW_IMG), dtype=torch.float32, device=self.device)
        self._check(self.lib.bw_expand_bits(self.handle, bits.data_ptr(), n, img.data_ptr()))
        return img

    # ------------------------------------------------------------------ read-back
    def get_state(self):
        E = self.num_envs
        blocks = np.zeros((E, L.BW_MAX_BLOCKS), dtype=self.dt["block"])
        n = np.zeros(E, dtype=np.int32)
        self._check(self.lib.bw_get_state(self.handle, blocks.ctypes.data, n.ctypes.data))
        return blocks, n

    def get_forces(self, variant=0):
        E = self.num_envs
        itf = np.zeros((E, L.BW_MAX_INTERFACES), dtype=self.dt["interface"])
        n = np.zeros(E, dtype=np.int32)
        self._check(self.lib.bw_get_forces(self.handle, variant, itf.ctypes.data, n.ctypes.data))
        return itf, n

    def lp_stats(self):
        """counters of the LP verdict path (all zero unless BW_LP_STATS was set when the handle was made)"""
        st = np.zeros(32, dtype=np.uint64)
        self._check(self.lib.bw_debug_lp_stats(self.handle, st.ctypes.data))
        return st

    def kernel_launches(self):
        return int(self.lib.bw_kernel_launches(self.handle))

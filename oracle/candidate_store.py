"""CPU restatement of the candidate store of the CUDA path (csrc/bw_actions.cu, `enumerate_store_kernel`): the
book-keeping that lets a step pose only the new block's candidates and test the others against the new pixels only.

TEST INFRASTRUCTURE (see oracle/__init__.py): this is not reference code -- the reference recomputes
generate_actions + get_action_features + filter_actions from scratch at every step
(robotoddler/training/successor_dqn.py:373-375, 403-406).  The restatement exists so that the store's rules can be
held against that from-scratch computation on the CPU (tests/test_oracle_candidate_store.py), independently of the GPU
parity tests:

  slot      one per possible candidate: (group = (shape, face), ground offset) or (group, target block, target face,
            offset).  It keeps the placement's products -- bounds flag (collision_on_action, gym_env.py:304-323) and
            raster (get_action_features) -- plus the overlap verdict of the last call that listed it and that call's stamp.
  drop      a block whose pose or shape differs from the copy taken when its slots were filled drops its slots.
  fresh     the block raster lost pixels or the obstacle raster changed since the last call (a reset): every listed
            candidate is tested against the whole raster.
  else      a candidate listed by the PREVIOUS call is tested against the pixels the block raster gained since then,
            and not at all once it overlaps; any other slot (stale stamp) is tested in full; a miss is posed,
            rasterised and tested in full.
"""
import collections

import numpy as np

from .actions import generate_actions
from .rendering import render_blocks_2d


class CandidateStore:
    def __init__(self, x_discr_ground, offset_values, xlim, ylim, img_size):
        self.ground = [float(v) for v in x_discr_ground]
        self.offsets = [float(v) for v in offset_values]
        self.xlim, self.ylim, self.img_size = xlim, ylim, img_size
        self.slots = {}                       # key -> dict(bad, img, ovl, stamp)
        self.kept = {}                        # block index -> (pose, shape name) its slots were filled for
        self.seen_block = np.zeros(img_size, dtype=bool)
        self.seen_obst = np.zeros(img_size, dtype=bool)
        self.call = 0
        self.stats = collections.Counter()

    def _key(self, gym, a):
        if a.target_block < 0:
            return (a.shape, a.face, -1, 0, self.ground.index(float(a.offset_x)))
        return (a.shape, a.face, a.target_block, a.target_face, self.offsets.index(float(a.offset_x)))

    def enumerate(self, gym, block_raster, obst_raster):
        """-> (actions, validity mask, rasters) of the current state, in generate_actions order."""
        block_raster = np.asarray(block_raster, dtype=bool)
        obst_raster = np.asarray(obst_raster, dtype=bool)
        fresh = bool((self.seen_block & ~block_raster).any() or (self.seen_obst != obst_raster).any())
        full = block_raster | obst_raster
        delta = full if fresh else (block_raster & ~self.seen_block)
        self.seen_block, self.seen_obst = block_raster.copy(), obst_raster.copy()
        prev, self.call = self.call, self.call + 1
        blocks = gym.assembly_env.blocks
        for bi, blk in enumerate(blocks):
            sig = (tuple(blk.pose), blk.name)
            if self.kept.get(bi) != sig:
                for key in [k for k in self.slots if k[2] == bi]:
                    del self.slots[key]
                self.kept[bi] = sig
                self.stats["blocks_dropped"] += 1
        actions = list(generate_actions(gym, self.ground, self.offsets))
        mask = np.zeros(len(actions), dtype=bool)
        rasters = []
        for i, a in enumerate(actions):
            key = self._key(gym, a)
            s = self.slots.get(key)
            if s is None:                                         # miss: pose, raster, bounds flag
                img = render_blocks_2d([gym.create_block(a)], xlim=self.xlim, ylim=self.ylim, img_size=self.img_size).astype(bool)
                s = self.slots[key] = dict(bad=bool(gym.collision_on_action(a, self.xlim, self.ylim)), img=img, ovl=False, stamp=-1)
                s["ovl"] = bool((img & full).any())
                self.stats["posed"] += 1
            elif not fresh and s["stamp"] == prev:                # listed by the previous call: new pixels only
                if not s["bad"] and not s["ovl"]:
                    s["ovl"] = bool((s["img"] & delta).any())
                    self.stats["incremental_tests"] += 1
                else:
                    self.stats["no_test"] += 1
            else:                                                 # fresh call / stale stamp: the whole raster
                s["ovl"] = bool((s["img"] & full).any())
                self.stats["full_tests"] += 1
            s["stamp"] = self.call
            mask[i] = not s["bad"] and not s["ovl"]
            rasters.append(s["img"])
        return actions, mask, rasters

"""Scratch: first verdict mismatch between the LP handle and the Newton handle on a bench-like rollout; dumps the
blocks of that environment for an oracle replay (GPU needed).  python tools/lp_debug_mismatch.py OUT.json"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bridges_b200.envs.batched import BatchedAssemblyGym
XG = [-2.0 + 2.0 * i / 9 for i in range(10)]
sq = 0.6
task = dict(obstacles=[(i * sq, 0, sq / 2) for i in range(1, 6)], targets=[(5 * sq + 2.5 * sq, 0, sq / 2)])
urdfs = ["shapes/trapezoid.urdf", "shapes/hexagon.urdf"]
E = 192
a = BatchedAssemblyGym(E, urdfs, max_steps=15)
os.environ["BW_NO_LP"] = "1"
b = BatchedAssemblyGym(E, urdfs, max_steps=15)
del os.environ["BW_NO_LP"]
a.reset(task); b.reset(task)
hist = [[] for _ in range(E)]
found = []
for k in range(70):
    a.enumerate_actions(XG, (0.0,), amax=1024, with_bits=False)
    acts, _ = a.select_random(seed=4242 + 17 * k)
    an = acts.cpu().numpy().view(a.dt["action"]).copy()
    a.step(acts); b.step(acts)
    oa, ob = a.read_out().copy(), b.read_out().copy()
    for e in range(E):
        hist[e].append([int(an[e]["target_block"]), int(an[e]["target_face"]), int(an[e]["shape"]), int(an[e]["face"]), float(an[e]["offset_x"]), float(an[e]["offset_y"])])
    bad = np.nonzero((oa["stable"] != ob["stable"]) | (oa["stable_unfrozen"] != ob["stable_unfrozen"]))[0]
    for e in bad:
        blocks, n = a.get_state()
        rec = dict(step=k, env=int(e), actions=hist[e][-int(n[e]):],
                   lp={f: (oa[e][f].tolist() if hasattr(oa[e][f], "tolist") else oa[e][f]) for f in ("stable", "stable_unfrozen", "residual", "residual_unfrozen", "solver_status", "lp_pivots", "newton_iters", "n_blocks", "n_interfaces")},
                   newton={f: (ob[e][f].tolist() if hasattr(ob[e][f], "tolist") else ob[e][f]) for f in ("stable", "stable_unfrozen", "residual", "residual_unfrozen", "solver_status", "newton_iters", "n_blocks", "n_interfaces")},
                   blocks=[[float(blocks[e][i][f]) for f in ("x", "z", "c", "s")] + [int(blocks[e][i]["shape"])] for i in range(int(n[e]))])
        found.append(rec)
        print(json.dumps(rec))
    if len(found) >= 3:
        break
    done = (oa["terminated"] | oa["truncated"]).astype(bool)
    for e in np.nonzero(done)[0]:
        hist[e] = []
    a.reset_done(); b.reset_done()
json.dump(found, open(sys.argv[1], "w"))

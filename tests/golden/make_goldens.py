"""Harvest golden vectors from the reference's stored notebook outputs.

Run in the build container (needs /root/reference):  python tests/golden/make_goldens.py
Writes tests/golden/notebook_goldens.json.  Sources (all outputs are stored in the .ipynb):
  * notebooks/AssemblyEnv.ipynb cell 21      -- horizontal_bridge_setup(7), mu=2.0, 8 scripted steps
  * notebooks/AssemblyEnv.ipynb cells 24-25  -- hard_tower_setup(), 10 scripted steps, full obs dicts
  * notebooks/CRA_Assembly.ipynb cell 6/7    -- three trapezoids: stable flags and interface count
  * notebooks/CRA_Assembly.ipynb cell 24     -- merged hexagon face dict (face order of the Action API)
  * notebooks/CRA_Assembly.ipynb cell 4      -- four equal compressions 0.75 (box 1x3x1 on a support)
  * notebooks/CRA_Assembly.ipynb cell 31     -- STL vertex welding order
"""
import ast
import json
import os
import re

REF = "/root/reference/notebooks"
HERE = os.path.dirname(os.path.abspath(__file__))


def cells(name):
    nb = json.load(open(os.path.join(REF, name)))
    out = {}
    for i, c in enumerate(nb["cells"]):
        if c["cell_type"] != "code":
            continue
        text = []
        for o in c.get("outputs", []):
            if "text" in o:
                text.append("".join(o["text"]))
            elif "data" in o and "text/plain" in o["data"]:
                text.append("".join(o["data"]["text/plain"]))
        out[i] = ("".join(c["source"]), "\n".join(text))
    return out


def parse_actions(src):
    acts = []
    for m in re.finditer(r"Action\(([^)]*)\)", src):
        args = m.group(1)
        if "target_block=" in args:
            kw = dict(re.findall(r"(\w+)=([-\w.]+)", args))
            acts.append([int(kw["target_block"]), int(kw["target_face"]), int(kw["shape"]), int(kw["face"]),
                         float(kw.get("offset_x", 0)), float(kw.get("offset_y", 0))])
        else:
            v = [float(x) for x in args.split(",")]
            acts.append([int(v[0]), int(v[1]), int(v[2]), int(v[3]), v[4], v[5]])
    return acts


def main():
    env_nb = cells("AssemblyEnv.ipynb")
    cra_nb = cells("CRA_Assembly.ipynb")
    gold = {}

    src, out = env_nb[21]
    steps = []
    for m in re.finditer(r"Stable: (\w+), Frozen Block: (\w+), Collision: (\w+), Targets Reached: (\d+)\n"
                         r"Reward: (-?\d+), Terminated: (\w+), Truncated: (\w+)", out):
        steps.append(dict(stable=m.group(1) == "True", collision=m.group(3) == "True", n_reached=int(m.group(4)),
                          reward=int(m.group(5)), terminated=m.group(6) == "True"))
    gold["horizontal_bridge_7_mu2"] = dict(source="AssemblyEnv.ipynb cell 21", mu=2.0, num_obstacles=7,
                                           actions=parse_actions(src), steps=steps)

    steps = []
    acts = parse_actions(env_nb[24][0]) + [a for a in parse_actions(env_nb[25][0])]
    # cell 25 has one commented-out step at the end
    acts = acts[:10]
    for text in (env_nb[24][1], env_nb[25][1]):
        for line in text.splitlines():
            if not line.startswith("({'blocks'"):
                continue
            clean = re.sub(r"Block \((\d+)\)", r"'Block\1'", line)
            obs, reward, terminated, truncated, info = ast.literal_eval(clean)
            steps.append(dict(stable=obs["stable"], collision=obs["collision"], reward=reward, terminated=terminated,
                              truncated=truncated, n_blocks=len(obs["blocks"]),
                              targets_remaining=obs["targets_remaining"], targets_reached=obs["targets_reached"],
                              distance_to_targets=obs["distance_to_targets"]))
    gold["hard_tower"] = dict(source="AssemblyEnv.ipynb cells 24-25", actions=acts, steps=steps)

    src, out = cra_nb[6]
    flags = re.findall(r"'stable': (\w+)", out)
    n_itf = int(re.search(r"Number of interfaces: (\d+)", cra_nb[7][1]).group(1))
    gold["three_trapezoids"] = dict(source="CRA_Assembly.ipynb cells 6-7",
                                    placements=[[-1, 0, 3, 0.35], [0, 1, 2, 0.8], [1, 1, 2, 0.0]],
                                    stable_after_block2_and_3=[f == "True" for f in flags], n_interfaces=n_itf)

    face_dict = ast.literal_eval(cra_nb[24][1].splitlines()[0])
    gold["hexagon_faces"] = dict(source="CRA_Assembly.ipynb cell 24", face={str(k): v for k, v in face_dict.items()})

    gold["box_on_support"] = dict(source="CRA_Assembly.ipynb cells 2-4",
                                  compressions=[float(x) for x in re.findall(r"Compression: ([\d.]+)", cra_nb[4][1])])

    verts = [ast.literal_eval(l) for l in cra_nb[31][1].splitlines() if l.startswith("[")]
    gold["trapezoid_txt_stl_vertices"] = dict(source="CRA_Assembly.ipynb cell 31", vertices=verts)

    with open(os.path.join(HERE, "notebook_goldens.json"), "w") as fh:
        json.dump(gold, fh, indent=1)
    print("wrote", len(gold), "golden groups")


if __name__ == "__main__":
    main()

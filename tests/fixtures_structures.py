"""Known-answer structures of the reference, re-encoded as data.

Source: assembly_gym/assembly_gym/utils/structures.py:22-108 -- each structure is a list of
(Action(target_block, target_face, shape, face, offset_x, offset_y, frozen), expected_stable)
with expected labels written as expressions of mu / freeze_last.  An action's `frozen` flag
says whether the block just placed is a support when the label is evaluated.
"""

TRAPEZOID, HEXAGON, CUBE = ["trapezoid"], ["hexagon"], ["cube"]


def hexagon(mu=0.8, freeze_last=True):
    # structures.py:22-30
    return TRAPEZOID, [((-1, 0, 0, 0, 0, 0, False), True), ((0, 3, 0, 3, 0., 0, False), mu > 1.732)]


def trapezoid_bridge(mu=0.8, freeze_last=True):
    # structures.py:33-48
    fl = freeze_last
    return TRAPEZOID, [
        ((-1, 0, 0, 0, -3, 0, fl), True),
        ((0, 3, 0, 3, 0., 0, fl), fl or mu > 1.732),
        ((1, 1, 0, 1, 0, 0, fl), fl and mu > 0.5),
        ((2, 3, 0, 3, 0, 0, fl), fl and mu > 0.5),
        ((3, 1, 0, 2, 0, 0, fl), fl and mu > 0.5),
        ((4, 0, 0, 1, 0, 0, fl), fl and mu > 0.5),
        ((5, 3, 0, 3, 0, 0, fl), fl and mu > 0.5),
        ((6, 1, 0, 1, 0, 0, fl), fl and mu > 0.5),
        ((7, 3, 0, 3, 0, 0, False), mu > 0.5)]


def hexagon_bridge_3(mu=0.8, freeze_last=True):
    # structures.py:50-59
    fl = freeze_last
    return HEXAGON, [((-1, 0, 0, 0, -3, 0, fl), True), ((0, 5, 0, 0, 0., 0, fl), fl), ((1, 5, 0, 0, 0., 0, False), fl)]


def hexagon_bridge_5(mu=0.8, freeze_last=True):
    # structures.py:61-71
    fl = freeze_last
    return HEXAGON, [((-1, 0, 0, 0, -3, 0, fl), True), ((0, 5, 0, 0, 0., 0, fl), fl), ((1, 4, 0, 0, 0., 0, fl), fl),
                     ((2, 5, 0, 0, 0., 0, fl), fl), ((3, 4, 0, 0, 0., 0, False), fl)]


def horizontal_bridge(mu=0.8, freeze_last=True):
    # structures.py:74-86
    fl = freeze_last
    return TRAPEZOID, [((-1, 0, 0, 2, -0.9, 0, fl), True), ((0, 0, 0, 2, 0, 0, fl), fl), ((1, 0, 0, 2, 0, 0, False), True)]


def tower(mu=0.8, freeze_last=True, num_blocks=3):
    # structures.py:89-98
    return CUBE, [((i - 1, 0, 0, 3, 0, 0, False), True) for i in range(num_blocks)]


def levitating_block(mu=0.8, freeze_last=False, offset_y=0.5):
    # structures.py:102-108
    fl = freeze_last
    return CUBE, [((-1, 0, 0, 0, 0, offset_y, fl), fl or offset_y < 1e-4), ((0, 3, 0, 0, 0, 0, fl), offset_y < 1e-4)]


STRUCTURES = dict(hexagon=hexagon, trapezoid_bridge=trapezoid_bridge, hexagon_bridge_3=hexagon_bridge_3,
                  hexagon_bridge_5=hexagon_bridge_5, horizontal_bridge=horizontal_bridge, tower=tower,
                  levitating_block=levitating_block)

# labels that are parameterised in mu (valid for every mu); the others only hold at the default mu = 0.8
MU_PARAMETERISED = ("hexagon", "trapezoid_bridge", "tower", "levitating_block")

# Known label/physics disagreements: the completed hexagon arches are
# physically stable with nothing frozen (need mu >= 0.577) while the label is the constant
# `freeze_last`.
KNOWN_LABEL_MISSES = {("hexagon_bridge_3", False, 2), ("hexagon_bridge_5", False, 4)}


def cases(mus=(0.8,)):
    for mu in mus:
        for name, fn in STRUCTURES.items():
            if mu != 0.8 and name not in MU_PARAMETERISED:
                continue
            for fl in (True, False):
                shapes, steps = fn(mu=mu, freeze_last=fl)
                yield name, mu, fl, shapes, steps

"""Block library: the oracle's restatement of compas' mesh pipeline against the reference's
stored notebook outputs, and the product's loader against the oracle's (CPU only)."""
import json
import os

import pytest

from bridges_b200.shapes_io import load_shape_tables
from oracle import compas_lite as cl
from oracle.assembly_env import DEFAULT_SHAPES_ROOT, Shape

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "notebook_goldens.json")))
NAMES = ["trapezoid", "hexagon", "cube", "cube1", "cube06", "block", "small_cube", "t_block", "v_block"]
REF_SHAPES = "/root/reference/assembly_gym/shapes"


def test_hexagon_face_dict_matches_notebook():
    # CRA_Assembly.ipynb cell 24: keys, key order and vertex cycles of the merged mesh
    mesh = cl.mesh_from_urdf(os.path.join(DEFAULT_SHAPES_ROOT, "shapes/hexagon.urdf"))
    cl.merge_coplanar_faces(mesh)
    want = {int(k): v for k, v in GOLD["hexagon_faces"]["face"].items()}
    assert list(mesh.face.keys()) == list(want.keys())
    assert mesh.face == want


def test_face_index_tables():
    # index = position among faces with |n_y| < 1e-6 in dict order
    t = Shape(urdf_file="shapes/trapezoid.urdf")
    assert t.num_faces_2d == 4
    signs = [(round(nx, 3), round(nz, 3)) for nx, nz in t.face_normals_2d]
    assert signs == [(-0.866, 0.5), (0.0, 1.0), (0.866, 0.5), (0.0, -1.0)]          # left, top, right, bottom
    h = Shape(urdf_file="shapes/hexagon.urdf")
    signs = [(round(nx, 3), round(nz, 3)) for nx, nz in h.face_normals_2d]
    assert signs == [(0.0, -1.0), (-0.866, -0.5), (0.866, -0.5), (-0.866, 0.5), (0.0, 1.0), (0.866, 0.5)]
    c = Shape(urdf_file="shapes/cube1.urdf")
    assert c.face_normals_2d == [(0.0, -1.0), (1.0, 0.0), (-1.0, 0.0), (0.0, 1.0)]  # bottom, +x, -x, top
    assert c.face_centers_2d == [(0.0, -0.5), (0.5, 0.0), (-0.5, 0.0), (0.0, 0.5)]
    # float32 STL coordinates widened to float64
    assert t.polygon_2d[0] == (1.0, -0.3595713675022125) and t.polygon_2d[2] == (-0.5, 0.5064539909362793)
    assert abs(t.area - 1.29903804) < 1e-8 and t.depth == 1.0


@pytest.mark.parametrize("name", NAMES)
def test_product_loader_equals_oracle_loader(name):
    o = Shape(urdf_file=f"shapes/{name}.urdf")
    t = load_shape_tables(f"shapes/{name}.urdf")
    assert t.normals == o.face_normals_2d
    assert t.centers == o.face_centers_2d
    assert [tuple(e) for e in t.ends] == [tuple(e) for e in o.face_ends_2d]
    assert t.polygon == o.polygon_2d
    assert (t.com, t.area, t.depth) == (o.centroid_2d, o.area, o.depth)
    assert t.aabb == o.bounding_box


@pytest.mark.skipif(not os.path.isdir(REF_SHAPES), reason="reference not mounted")
@pytest.mark.parametrize("name", NAMES)
def test_shipped_library_equals_reference_files(name):
    ours = load_shape_tables(f"shapes/{name}.urdf")
    ref = load_shape_tables(os.path.join(REF_SHAPES, f"{name}.urdf"))
    for attr in ("normals", "centers", "ends", "polygon", "com", "area", "depth", "aabb", "vertices3d"):
        assert getattr(ours, attr) == getattr(ref, attr), attr
    if name in ("trapezoid", "hexagon", "t_block", "v_block"):
        a = open(os.path.join(DEFAULT_SHAPES_ROOT, "shapes/blocks", f"{name}.stl"), "rb").read()
        b = open(os.path.join(REF_SHAPES, "blocks", f"{name}.stl"), "rb").read()
        assert a == b


@pytest.mark.skipif(not os.path.isdir(REF_SHAPES), reason="reference not mounted")
def test_stl_weld_order_matches_notebook():
    # CRA_Assembly.ipynb cell 31 printed the welded vertices of the (then float32) trapezoid
    mesh = cl.mesh_from_stl(os.path.join(REF_SHAPES, "blocks", "trapezoid-rescaled-txt.stl"))
    got = [[round(c, 6) for c in mesh.vertex[k]] for k in mesh.vertices()]
    want = [[round(c, 6) for c in v] for v in GOLD["trapezoid_txt_stl_vertices"]["vertices"]]
    assert got == want


def test_missing_urdf_raises_like_reference():
    with pytest.raises(FileNotFoundError):
        Shape(urdf_file="shapes/does_not_exist.urdf")
    with pytest.raises(FileNotFoundError):
        load_shape_tables("shapes/does_not_exist.urdf")

"""Named parity cases shared by the golden generator and the CPU/GPU tests."""
from oracle import binding as ob
from realtrace_b200 import scenes

SMALL = (160, 120)

GOLDEN_CASES = [
    "tetra_d10", "bob2000_d10", "analytic_close_d5", "analytic_notetra_d5", "analytic_stock_d1",
    "bobtex_d3", "blubmixed_d5", "synth_small_d1", "bob_full_bboxfixed_d10",
]


def build_case(name):
    """-> (scene, camera, max_depth, oracle mode)"""
    w, h = SMALL
    if name == "tetra_d10":
        return scenes.obj_scene("tetrahedron.obj"), scenes.stock_camera(w, h), 10, ob.MODE_AS_SHIPPED
    if name == "bob2000_d10":
        return scenes.obj_scene("bob_tri.obj", 2000), scenes.stock_camera(w, h), 10, ob.MODE_AS_SHIPPED
    if name == "analytic_close_d5":
        return scenes.analytic_scene(), scenes.close_camera(w, h), 5, ob.MODE_AS_SHIPPED
    if name == "analytic_notetra_d5":
        return scenes.analytic_scene(with_tetrahedron=False), scenes.close_camera(w, h), 5, ob.MODE_AS_SHIPPED
    if name == "analytic_stock_d1":
        return scenes.analytic_scene(), scenes.stock_camera(w, h), 1, ob.MODE_AS_SHIPPED
    if name == "bobtex_d3":
        return scenes.bob_textured(), scenes.stock_camera(w, h), 3, ob.MODE_AS_SHIPPED
    if name == "blubmixed_d5":
        return scenes.blub_mixed(), scenes.stock_camera(w, h), 5, ob.MODE_AS_SHIPPED
    if name == "synth_small_d1":
        return (scenes.synthetic_sphere_grid(grid=3, level=2), scenes.synthetic_camera(w, h, grid=3), 1,
                ob.MODE_AS_SHIPPED)
    if name == "bob_full_bboxfixed_d10":
        return scenes.obj_scene("bob_tri.obj"), scenes.stock_camera(w, h), 10, ob.MODE_BBOX_FIXED
    raise KeyError(name)

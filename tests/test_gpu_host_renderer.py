"""The C++ host `Renderer` / `GSScene` (3dgs.cpp_b200/host) driven through the vkgs_*-style bridge:
load .ply -> set camera -> render(width, height) -> RGBA buffer, against the oracle."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ply(gs, tmp_path_factory):
    rec = gs.synth_records(42, 10_000)
    path = tmp_path_factory.mktemp("scene") / "c1.ply"
    gs.write_ply(path, rec)
    return path, gs.activate_records(rec)


def oracle_image(oracle, vtx, u, mode=1):
    oracle.set_exp_mode(mode)
    try:
        return oracle.render_frame(vtx, oracle.cov3d(vtx), u)["rgba"]
    finally:
        oracle.set_exp_mode(0)


def test_renderer_render_matches_oracle(gs, oracle, ply):
    path, vtx = ply
    r = gs.HostRenderer(path, device=0, width=640, height=480, fmt=gs.FORMAT_BGRA8)
    try:
        assert r.num_vertices == 10_000
        # default camera of the reference: origin, identity rotation, fov 45 (Renderer.h:79-85)
        pos, quat, fov = r.get_camera()
        assert list(pos) == [0, 0, 0] and list(quat) == [1, 0, 0, 0] and fov == 45.0
        r.set_camera([0, 0, 5], [1, 0, 0, 0])
        img = r.render(640, 480, gs.FORMAT_RGBA32F)
        u = gs.uniforms_from_camera([0, 0, 5], [1, 0, 0, 0], 45.0, 0.1, 1000.0, 640, 480)
        ref = oracle_image(oracle, vtx, u)
        assert np.array_equal(img, ref)
        # draw(): configured size, swapchain format B8G8R8A8_UNORM
        bgra = r.draw()
        assert np.array_equal(bgra, oracle.pack_unorm8(ref, bgra=True))
        st = r.stats()
        assert st.num_gaussians == 10_000 and st.num_instances > 0 and st.frame_ms > 0
    finally:
        r.close()


def test_renderer_camera_controls_follow_the_reference(gs, oracle, ply):
    path, vtx = ply
    r = gs.HostRenderer(path, device=0, width=320, height=240, fmt=gs.FORMAT_RGBA8)
    try:
        q = scenes.quat_axis_angle([0, 1, 0], 20)
        r.set_camera([0.5, 0.0, 6.0], q)
        r.movement(0.1, -0.2, 0.3)                       # vkgs_movement -> Camera::translate
        pos, quat, _ = r.get_camera()
        assert np.array_equal(pos, oracle.camera_translate([0.5, 0.0, 6.0], q, [0.1, -0.2, 0.3]))
        r.keys([1, 0, 0, 0, 0, 0])                       # W: forward 0.3 along -z of the camera
        r.pan(10.0, -4.0)                                # cursor delta -> two quaternion rotations
        pos2, quat2, _ = r.get_camera()
        assert not np.array_equal(pos2, pos) and not np.array_equal(quat2, quat)
        assert abs(np.linalg.norm(quat2) - 1.0) < 1e-5
        img = r.render(320, 240, gs.FORMAT_RGBA32F)
        u = gs.uniforms_from_camera(pos2, quat2, 45.0, 0.1, 1000.0, 320, 240)
        assert np.array_equal(img, oracle_image(oracle, vtx, u))
    finally:
        r.close()


def test_renderer_missing_scene_raises(gs):
    with pytest.raises(RuntimeError, match="File does not exist"):
        gs.HostRenderer("/nonexistent.ply")

"""C++ host side (GSScene loader, Renderer::updateUniforms, camera, PLY I/O, synthetic scenes) against
the oracle's independent C restatement.  No GPU needed."""
import numpy as np
import pytest

import scenes


def test_activation_matches_oracle_bit_exact(gs, oracle):
    rec = gs.synth_records(11, 5000)
    assert np.array_equal(gs.activate_records(rec), oracle.load_records(rec))


def test_sh_interleave_is_channel_major_to_rgb(gs):
    rec = np.zeros((1, 62), np.float32)
    rec[0, 6:9] = [100, 101, 102]            # f_dc
    rec[0, 9:54] = np.arange(45)             # f_rest_k = k : 15 R, then 15 G, then 15 B
    rec[0, 58] = 1.0                         # rot w
    v = gs.activate_records(rec)[0]
    sh = v[12:]
    assert list(sh[:3]) == [100, 101, 102]
    for j in range(1, 16):
        for c in range(3):
            assert sh[3 * j + c] == 15 * c + (j - 1)   # GSScene.cpp:51-55
    assert v[3] == 1.0 and np.allclose(v[4:7], 1.0) and v[7] == 0.5 and list(v[8:12]) == [1, 0, 0, 0]


def test_uniforms_match_oracle_bit_exact_over_random_cameras(gs, oracle):
    rng = np.random.default_rng(5)
    for k in range(200):
        pos = rng.uniform(-10, 10, 3)
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        fov = float(rng.uniform(20, 100))
        w, h = int(rng.integers(16, 4000)), int(rng.integers(16, 2500))
        a = gs.uniforms_from_camera(pos, q, fov, 0.1, 1000.0, w, h)
        b = oracle.uniforms_from_camera(pos, q, fov, 0.1, 1000.0, w, h)
        assert bytes(a) == bytes(b), k


def test_uniforms_geometry(gs):
    """view maps the camera position to the origin; camera space is x right, y down, z forward."""
    pos = np.array([1.0, 2.0, 3.0], np.float32)
    q = scenes.quat_axis_angle([0.2, 1.0, 0.1], 33)
    u = gs.uniforms_from_camera(pos, q, 50.0, 0.1, 1000.0, 800, 600)
    V = np.array(u.view_mat, np.float64).reshape(4, 4).T
    P = np.array(u.proj_mat, np.float64).reshape(4, 4).T
    assert np.allclose(V @ np.append(pos, 1.0), [0, 0, 0, 1], atol=1e-5)
    R = V[:3, :3]
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-5)
    assert np.isclose(np.linalg.det(R), 1.0, atol=1e-5)  # two row flips keep det +1
    # a point 4 units in front of the camera (camera looks down -z in its own frame)
    w = np.array([0, 0, -4.0])
    # rotate by q
    qw, qx, qy, qz = q.astype(np.float64)
    Rq = np.array([[1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qw * qz), 2 * (qx * qz + qw * qy)],
                   [2 * (qx * qy + qw * qz), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qw * qx)],
                   [2 * (qx * qz - qw * qy), 2 * (qy * qz + qw * qx), 1 - 2 * (qx * qx + qy * qy)]])
    pw = np.append(Rq @ w + pos, 1.0)
    pv = V @ pw
    assert np.allclose(pv[:3], [0, 0, 4.0], atol=1e-4)
    ph = P @ pw
    assert np.isclose(ph[3], 4.0, atol=1e-4)              # p_hom.w = z_c
    assert np.allclose(ph[:2] / ph[3], [0, 0], atol=1e-5)  # optical axis -> ndc centre
    assert np.isclose(u.tan_fovx, np.tan(np.radians(50.0) / 2), rtol=1e-6)
    assert np.isclose(u.tan_fovy, u.tan_fovx * 600 / 800, rtol=1e-6)
    # +x world-right of camera lands at +ndc.x; camera-up lands at -ndc.y (y down)
    pr = np.append(Rq @ np.array([1.0, 0.5, -4.0]) + pos, 1.0)
    nd = (P @ pr)[:2] / (P @ pr)[3]
    assert nd[0] > 0 and nd[1] < 0


def test_camera_translate_matches_oracle(gs, oracle):
    rng = np.random.default_rng(9)
    for _ in range(50):
        pos, t = rng.uniform(-5, 5, 3), rng.uniform(-1, 1, 3)
        q = rng.normal(size=4)
        q /= np.linalg.norm(q)
        assert np.array_equal(gs.camera_translate(pos, q, t), oracle.camera_translate(pos, q, t))


def test_ply_roundtrip_and_both_loaders_agree(gs, oracle, tmp_path):
    rec = gs.synth_records(21, 1234)
    path = tmp_path / "scene.ply"
    gs.write_ply(path, rec)
    head = path.read_bytes()[:2000]
    assert head.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex 1234\n") and b"property float f_rest_44" in head
    a = gs.load_ply(path)
    b = oracle.load_ply(path)
    assert a.shape == (1234, 60)
    assert np.array_equal(a, b)
    assert np.array_equal(a, gs.activate_records(rec))


def test_large_ply_is_read_by_several_threads_and_stays_exact(gs, tmp_path):
    """GSScene::loadToHost splits the body over threads (pread on disjoint slices) once there are more than 65536 records:
    the result must equal the single-pass activation, record for record, and a file cut inside a later slice is an error."""
    n = 200_003  # 4 slices, the last one ragged
    rec = gs.synth_records(7, n)
    path = tmp_path / "big.ply"
    gs.write_ply(path, rec)
    got = gs.load_ply(path)
    assert got.shape == (n, 60) and np.array_equal(got, gs.activate_records(rec))
    data = path.read_bytes()
    path.write_bytes(data[:-5000])
    with pytest.raises(RuntimeError, match="Unexpected end of file"):
        gs.load_ply(path)


def test_missing_scene_file_raises_like_the_reference(gs):
    with pytest.raises(RuntimeError, match="File does not exist"):
        gs.load_ply("/nonexistent/scene.ply")


def test_truncated_ply_is_an_error(gs, tmp_path):
    rec = gs.synth_records(1, 10)
    path = tmp_path / "t.ply"
    gs.write_ply(path, rec)
    data = path.read_bytes()
    path.write_bytes(data[:-100])
    with pytest.raises(RuntimeError):
        gs.load_ply(path)


CANON = (["x", "y", "z", "nx", "ny", "nz", "f_dc_0", "f_dc_1", "f_dc_2"] + [f"f_rest_{k}" for k in range(45)]
         + ["opacity", "scale_0", "scale_1", "scale_2", "rot_0", "rot_1", "rot_2", "rot_3"])


def _write_custom_ply(path, fields, n, fmt="binary_little_endian"):
    """fields: list of (name, ply type, numpy dtype, column array)."""
    dt = np.dtype([(name, npt) for name, _, npt, _ in fields])
    arr = np.zeros(n, dtype=dt)
    for name, _, _, col in fields:
        arr[name] = col
    head = f"ply\nformat {fmt} 1.0\ncomment custom layout\nelement vertex {n}\n"
    head += "".join(f"property {plyt} {name}\n" for name, plyt, _, _ in fields) + "end_header\n"
    path.write_bytes(head.encode() + arr.tobytes())


def test_ply_header_property_list_is_honoured(gs, tmp_path):
    """shuffled order, extra properties, a double field, SH degree 1 (9 f_rest, channel-major)."""
    n = 321
    rec = gs.synth_records(5, n)
    K = 3  # degree 1: 3 coefficients per channel
    rest = rec[:, 9:54].reshape(n, 3, 15)
    want = rec.copy()
    want[:, 3:6] = 0.0  # normals are not stored in this file
    w_rest = np.zeros((n, 3, 15), np.float32)
    w_rest[:, :, :K] = rest[:, :, :K]
    want[:, 9:54] = w_rest.reshape(n, 45)
    col = {name: rec[:, k] for k, name in enumerate(CANON)}
    fields = [("rot_3", "float", "<f4", col["rot_3"]), ("label", "uchar", "u1", np.arange(n) % 250),
              ("opacity", "double", "<f8", col["opacity"].astype(np.float64))]
    fields += [(f"f_rest_{c * K + j}", "float", "<f4", rest[:, c, j]) for c in range(3) for j in range(K)]
    fields += [(nm, "float32", "<f4", col[nm]) for nm in ("z", "y", "x", "scale_2", "scale_0", "scale_1", "rot_0", "rot_1", "rot_2")]
    fields += [("id", "int", "<i4", np.arange(n)), ("f_dc_1", "float", "<f4", col["f_dc_1"]), ("f_dc_0", "float", "<f4", col["f_dc_0"]),
               ("f_dc_2", "float", "<f4", col["f_dc_2"])]
    path = tmp_path / "custom.ply"
    _write_custom_ply(path, fields, n)
    got = gs.load_ply(path)
    assert np.array_equal(got, gs.activate_records(want))


def test_ply_without_property_list_is_read_like_the_reference(gs, tmp_path):
    rec = gs.synth_records(6, 17)
    path = tmp_path / "bare.ply"
    path.write_bytes(b"ply\nformat binary_little_endian 1.0\nelement vertex 17\nend_header\n" + rec.tobytes())
    assert np.array_equal(gs.load_ply(path), gs.activate_records(rec))


@pytest.mark.parametrize("case", ["missing_opacity", "ascii", "bad_rest_count", "int_position", "list_property"])
def test_malformed_ply_headers_are_errors(gs, tmp_path, case):
    n = 4
    rec = gs.synth_records(7, n)
    col = {name: rec[:, k] for k, name in enumerate(CANON)}
    names = [nm for nm in CANON if not nm.startswith("n")]
    fields = [(nm, "float", "<f4", col[nm]) for nm in names]
    fmt = "binary_little_endian"
    path = tmp_path / f"{case}.ply"
    if case == "missing_opacity":
        fields = [f for f in fields if f[0] != "opacity"]
    elif case == "ascii":
        fmt = "ascii"
    elif case == "bad_rest_count":
        fields = [f for f in fields if f[0] != "f_rest_44"]
    elif case == "int_position":
        fields = [("x", "int", "<i4", np.zeros(n))] + [f for f in fields if f[0] != "x"]
    _write_custom_ply(path, fields, n, fmt)
    if case == "list_property":
        data = path.read_bytes().replace(b"end_header\n", b"property list uchar int vertex_indices\nend_header\n", 1)
        path.write_bytes(data)
    with pytest.raises(RuntimeError):
        gs.load_ply(path)


def test_synth_is_counter_based(gs):
    whole = gs.synth_records(42, 3000)
    parts = np.concatenate([gs.synth_records(42, 1000, first=0), gs.synth_records(42, 2000, first=1000)])
    assert np.array_equal(whole, parts)
    assert not np.array_equal(whole, gs.synth_records(43, 3000))
    assert np.all(whole[:, 3:6] == 0)                          # normals
    assert np.all(np.abs(whole[:, :3]) <= 3.0)                 # config-1 box
    assert whole[:, 54].min() >= -2 and whole[:, 54].max() <= 4  # opacity logits


def test_band_partition(gs):
    for h in (16, 480, 1400, 2160, 4320):
        for world in (1, 2, 3, 4, 8):
            tiles_y = (h + 15) // 16
            covered = []
            for r in range(world):
                b, e, per = gs.band_for_rank(h, r, world)
                assert 0 <= b <= e <= tiles_y and e - b <= per
                covered += list(range(b, e))
            assert covered == list(range(tiles_y))

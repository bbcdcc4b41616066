"""The tiny-input unit cases of test_oracle_units.py, on the CUDA library (compared with the CPU oracle), plus ABI
error behaviour."""
import numpy as np
import pytest

import parity as P
from conftest import sha as P_sha
import test_oracle_units as U
from forkerrenderer_b200 import binding as B

pytestmark = pytest.mark.gpu


def both(gpu_fgl, oracle_fgl, fn):
    fn(gpu_fgl)
    fn(oracle_fgl)


def test_edge_rules_and_ties(gpu_fgl, oracle_fgl):
    cases = [
        ([([(2, 2), (10, 2), (2, 10)], 0.5)], (16, 16)),
        ([([(1, 1), (12, 1), (1, 12)], 0.5), ([(1, 1), (12, 1), (1, 12)], 0.5), ([(3, 3), (9, 3), (3, 9)], 0.25)], (16, 16)),
        ([([(3, 3), (3, 3), (9, 9)], 0.5), ([(-40, -40), (-30, -40), (-40, -30)], 0.5), ([(2, 2), (6, 6), (10, 10)], 0.5)], (16, 16)),
        ([([(-100, -100), (300, -100), (-100, 300)], 0.5)], (16, 12)),
        ([([(0, 0), (63, 0), (0, 40)], 0.5), ([(63, 40), (63, 0), (0, 40)], 0.5), ([(5, 5), (70, 20), (30, 90)], 0.3)], (64, 41)),
        ([([(-5000000, -3), (5000000, 7), (11, 4000000)], 0.7), ([(1, 1), (30, 2), (2, 30)], 0.6)], (33, 31)),   # beyond the exact integer pre-filter
    ]
    for tris, (W, H) in cases:
        for f in (gpu_fgl, oracle_fgl):
            U.ndc_tri_scene(f, tris, W, H)
        for plane in ("ids_light", "depth", "shadow"):
            assert P.bits_equal(gpu_fgl.read_plane(plane), oracle_fgl.read_plane(plane)), (plane, W, H)


def test_in_place_gaussian(gpu_fgl):
    for (W, H) in ((13, 9), (300, 70), (1, 1), (3, 2), (129, 33)):
        rng = np.random.RandomState(W * 31 + H)
        a = rng.rand(H, W).astype(np.float32)
        gpu_fgl.init_geometry_buffers(W, H)
        gpu_fgl.write_plane("ao", a)
        gpu_fgl.blur(B.PLANE_AO, B.BLUR_TWO_PASS_GAUSSIAN)
        got = gpu_fgl.read_plane("ao")
        if W * H <= 4000:
            assert np.array_equal(got, U.blur_reference(a)), (W, H)
        else:
            orc = B.Fgl(P.ORACLE_LIB)
            orc.init_geometry_buffers(W, H)
            orc.write_plane("ao", a)
            orc.blur(B.PLANE_AO, B.BLUR_TWO_PASS_GAUSSIAN)
            assert np.array_equal(got, orc.read_plane("ao")), (W, H)
            orc.close()


def test_three_channel_blur(gpu_fgl, oracle_fgl):
    rng = np.random.RandomState(11)
    a = rng.rand(40, 50, 3).astype(np.float32)
    for f in (gpu_fgl, oracle_fgl):
        f.init_geometry_buffers(50, 40)
        f.write_plane("albedo", a)
        f.blur(B.PLANE_ALBEDO, B.BLUR_TWO_PASS_GAUSSIAN)
    assert np.array_equal(gpu_fgl.read_plane("albedo"), oracle_fgl.read_plane("albedo"))


@pytest.mark.parametrize("kind", sorted(P.BUFFER_KINDS))
def test_buffer_post_processing_matches_reference(kind, gpu_fgl, golden):
    """Both Buffer blurs on the device against the reference's own Buffer classes (tests/golden, `buffer_ops`)."""
    for W, H in P.BUFFER_SHAPES:
        a1, a3 = P.buffer_test_inputs(W, H)
        g1, g3 = P.blur_through_abi(gpu_fgl, P.BUFFER_KINDS[kind], a1, a3)
        want = golden["buffer_ops"]["%s_%dx%d" % (kind, W, H)]
        assert P_sha(g1) == want["buffer1f"]["sha256"], (kind, W, H, "Buffer1f")
        assert P_sha(g3) == want["buffer3f"]["sha256"], (kind, W, H, "Buffer3f")


def test_simple_blur_wavefront_across_passes(gpu_fgl, oracle_fgl):
    """The in-place 3x3 box is a wavefront over the whole plane; more than 1024 rows take several kernel passes."""
    rng = np.random.RandomState(23)
    for (W, H) in ((300, 1100), (7, 2050), (1500, 3)):
        a1, a3 = rng.rand(H, W).astype(np.float32), rng.rand(H, W, 3).astype(np.float32)
        g = P.blur_through_abi(gpu_fgl, B.BLUR_SIMPLE_3X3, a1, a3)
        o = P.blur_through_abi(oracle_fgl, B.BLUR_SIMPLE_3X3, a1, a3)
        assert np.array_equal(g[0], o[0]) and np.array_equal(g[1], o[1]), (W, H)


def test_ssaa_box(gpu_fgl, oracle_fgl):
    rng = np.random.RandomState(5)
    for (W, H, k) in ((8, 6, 2), (30, 21, 3), (64, 64, 4)):
        frame = rng.rand(H, W, 3).astype(np.float32)
        for f in (gpu_fgl, oracle_fgl):
            f.init_frame_buffer(W, H)
            f.write_plane("frame", frame)
            f.ssaa_resolve(k)
        assert np.array_equal(gpu_fgl.read_plane("ssaa_u8"), oracle_fgl.read_plane("ssaa_u8"))
        assert np.array_equal(gpu_fgl.read_plane("frame_u8"), oracle_fgl.read_plane("frame_u8"))


def test_plane_roundtrip_and_clears(gpu_fgl):
    gpu_fgl.init_geometry_buffers(17, 5)
    gpu_fgl.init_depth_buffer(17, 5)
    gpu_fgl.init_frame_buffer(17, 5)
    assert np.all(gpu_fgl.read_plane("ao") == 1.0) and np.all(gpu_fgl.read_plane("normal") == 0.0)
    assert np.all(gpu_fgl.read_plane("depth") == np.finfo(np.float32).max)
    gpu_fgl.clear_color((0.12, 0.5, 0.25))
    fr = gpu_fgl.read_plane("frame")
    assert np.allclose(fr[..., 0], 0.12) and np.allclose(fr[..., 1], 0.5) and np.allclose(fr[..., 2], 0.25)
    a = np.random.RandomState(1).rand(5, 17, 3).astype(np.float32)
    gpu_fgl.write_plane("worldpos", a)
    assert np.array_equal(gpu_fgl.read_plane("worldpos"), a)


def test_errors_are_reported_not_fatal(gpu_fgl):
    f = gpu_fgl
    with pytest.raises(B.FglError):
        f.read_plane(B.PLANE_FRAME_RGB8)                       # no frame buffer yet
    with pytest.raises(B.FglError):
        f.draw_mesh(99, B.SHADER_G, B.FglUniforms())           # bad handle
    v = f.upload_vertices(np.zeros((3, 3), np.float32), np.zeros((1, 2), np.float32), np.array([[0, 0, 1]], np.float32))
    with pytest.raises(B.FglError):
        f.upload_mesh(v, np.array([[0, 1, 7]], np.int32), np.zeros((1, 3), np.int32), np.zeros((1, 3), np.int32))   # index out of range
    m = f.upload_mesh(v, np.array([[0, 1, 2]], np.int32), np.zeros((1, 3), np.int32), np.zeros((1, 3), np.int32))
    f.init_geometry_buffers(8, 8)
    f.init_depth_buffer(8, 8)
    f.set_pass_type(B.PASS_GEOMETRY)
    with pytest.raises(B.FglError):
        f.draw_mesh(m, B.SHADER_BLINN_PHONG, B.FglUniforms())  # the reference's dynamic_cast<GShader&> would throw
    f.draw_mesh(m, B.SHADER_G, B.FglUniforms())                # and the context is still usable
    assert f.read_plane("depth").shape == (8, 8)

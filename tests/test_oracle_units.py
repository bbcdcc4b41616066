"""Unit behaviours of the restated path on tiny inputs (the reference has no tests; these are the cases SURVEY.md §4
asks for): inclusive edges on integer-snapped vertices, first-submitted-wins depth ties, bounding-box clamping of
unclipped geometry, the in-place blur recurrence, the SSAA integer box, texture wrap / filter modes."""
import numpy as np
import pytest

from forkerrenderer_b200 import binding as B


def ndc_tri_scene(f, tris_px, W, H, depth=0.5):
    """Draws screen-space triangles (pixel coordinates) through the shadow pass with identity matrices."""
    f.set_shadow_status(1)
    f.begin_frame()
    f.set_viewport(0, 0, W, H)
    f.init_shadow_buffer(W, H)
    f.init_depth_buffer(W, H)
    f.set_pass_type(B.PASS_SHADOW)
    eye = np.eye(4, dtype=np.float32)
    for tri, z in tris_px:
        pos = np.array([[(x + 0.5) / W * 2 - 1, (y + 0.5) / H * 2 - 1, z * 2 - 1] for x, y in tri], dtype=np.float32)
        v = f.upload_vertices(pos, np.zeros((1, 2), np.float32), np.array([[0, 0, 1]], np.float32))
        idx = np.array([[0, 1, 2]], np.int32)
        m = f.upload_mesh(v, idx, np.zeros_like(idx), np.zeros_like(idx))
        f.draw_mesh(m, B.SHADER_DEPTH, B.FglUniforms(model=eye, light_space=eye))


def test_edges_are_inclusive_and_vertices_snap(oracle_fgl):
    f = oracle_fgl
    ndc_tri_scene(f, [([(2, 2), (10, 2), (2, 10)], 0.5)], 16, 16)
    ids = f.read_plane("ids_light")
    cov = ids >= 0
    assert cov[2, 2] and cov[2, 10] and cov[10, 2]          # the three vertices themselves
    assert cov[2, 5] and cov[5, 2] and cov[6, 6]            # all three edges, hypotenuse x + y = 12 included
    assert not cov[7, 6] and not cov[1, 2] and not cov[2, 1]
    assert cov.sum() == 45                                  # 9 + 8 + ... + 1 lattice points


def test_first_submitted_triangle_wins_depth_ties(oracle_fgl):
    f = oracle_fgl
    tri = [(1, 1), (12, 1), (1, 12)]
    ndc_tri_scene(f, [(tri, 0.5), (tri, 0.5), (tri, 0.25)], 16, 16)
    ids = f.read_plane("ids_light")
    assert set(np.unique(ids)) == {-1, 2}                   # the nearer third triangle wins everywhere it covers
    f2 = B.Fgl(f.lib)
    ndc_tri_scene(f2, [(tri, 0.5), (tri, 0.5)], 16, 16)
    assert set(np.unique(f2.read_plane("ids_light"))) == {-1, 0}   # equal depth: strict-less test keeps the first
    f2.close()


def test_degenerate_and_offscreen_triangles_cover_nothing(oracle_fgl):
    f = oracle_fgl
    ndc_tri_scene(f, [([(3, 3), (3, 3), (9, 9)], 0.5), ([(-40, -40), (-30, -40), (-40, -30)], 0.5), ([(2, 2), (6, 6), (10, 10)], 0.5)], 16, 16)
    assert (f.read_plane("ids_light") >= 0).sum() == 0


def test_unclipped_triangle_is_clamped_to_the_buffer(oracle_fgl):
    f = oracle_fgl
    ndc_tri_scene(f, [([(-100, -100), (300, -100), (-100, 300)], 0.5)], 16, 12)
    assert (f.read_plane("ids_light") == 0).all()
    d = f.read_plane("depth")
    assert np.all(np.abs(d - 0.5) < 1e-6)


def blur_reference(a):
    g = np.array([0.227027, 0.1945946, 0.1216216, 0.054054, 0.016216], dtype=np.float32)
    a = a.copy()
    H, W = a.shape
    for h in range(H):
        for w in range(W):
            r = np.float32(a[h, w] * g[0])
            for i in range(1, 5):
                r = np.float32(r + np.float32(a[h, min(w + i, W - 1)] * g[i]))
                r = np.float32(r + np.float32(a[h, max(w - i, 0)] * g[i]))
            a[h, w] = r
    for h in range(H):
        for w in range(W):
            r = np.float32(a[h, w] * g[0])
            for i in range(1, 5):
                r = np.float32(r + np.float32(a[min(h + i, H - 1), w] * g[i]))
                r = np.float32(r + np.float32(a[max(h - i, 0), w] * g[i]))
            a[h, w] = r
    return a


def test_in_place_gaussian_is_a_recurrence(oracle_fgl):
    f = oracle_fgl
    rng = np.random.RandomState(3)
    a = rng.rand(9, 13).astype(np.float32)
    f.init_geometry_buffers(13, 9)
    f.write_plane("ao", a)
    f.blur(B.PLANE_AO, B.BLUR_TWO_PASS_GAUSSIAN)
    got = f.read_plane("ao")
    assert np.array_equal(got, blur_reference(a))
    # and it is NOT the symmetric convolution of the original values
    assert np.abs(got - a).max() > 1e-3


def test_ssaa_integer_box(oracle_fgl):
    f = oracle_fgl
    rng = np.random.RandomState(5)
    frame = rng.rand(6, 8, 3).astype(np.float32)
    f.init_frame_buffer(8, 6)
    f.write_plane("frame", frame)
    f.ssaa_resolve(2)
    q = (frame * np.float32(254.99)).astype(np.uint8).astype(np.int64)
    want = (q.reshape(3, 2, 4, 2, 3).sum(axis=(1, 3)) / 4.0).astype(np.int64).astype(np.uint8)
    assert np.array_equal(f.read_plane("ssaa_u8"), want)
    assert np.array_equal(f.read_plane("frame_u8"), q.astype(np.uint8))


def test_host_alloc_backs_plane_reads(oracle_fgl):
    """fgl_host_alloc / fgl_host_free (page-locked memory in the product, malloc in the oracle): read_plane(pinned=True)
    returns the same data as a pageable read and reuses its buffer."""
    rng = np.random.RandomState(3)
    a = rng.rand(6, 9).astype(np.float32)
    oracle_fgl.init_geometry_buffers(9, 6)
    oracle_fgl.write_plane("ao", a)
    p1 = oracle_fgl.read_plane("ao", pinned=True)
    assert np.array_equal(p1, a) and np.array_equal(oracle_fgl.read_plane("ao"), a)
    oracle_fgl.write_plane("ao", a * 2)
    p2 = oracle_fgl.read_plane("ao", pinned=True)
    assert p2.ctypes.data == p1.ctypes.data and np.array_equal(p2, a * 2)

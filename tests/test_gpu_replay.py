"""Multi-frame use (SURVEY.md §8 f, N4): a frame recorded as a CUDA graph (fgl_frame_record_* / fgl_frame_replay, frh_render_replay)
must be the frame the same calls render eagerly — bit for bit, every plane — on every replay, after the camera or the light
has moved (the graph is patched in place), and frames that cannot be recorded must fall back to eager rendering, loudly
described, never to a wrong image.  The eager frames themselves are pinned to the reference by tests/test_gpu_scenes.py."""
import os

import numpy as np
import pytest

import parity as P
from conftest import EXACT_PLANES, sha
from forkerrenderer_b200 import binding as B
from forkerrenderer_b200.synthetic import SyntheticScene

pytestmark = pytest.mark.gpu

PLANES = ["depth", "shadow", "frame", "frame_u8", "ids_camera", "ids_light", "normal", "worldpos", "lightndc", "albedo", "param", "ao"]


def read_all(host, sc):
    return {n: host.fgl.read_plane(n).copy() for n in PLANES + (["ssaa_u8"] if sc.ssaa else [])}


def assert_same(a, b, what):
    for n in a:
        assert P.bits_equal(a[n], b[n]), "%s: plane %s of the replayed frame differs from the eager frame" % (what, n)


# (scene, shadow filter): hard shadows; PCF (disk-sample table); SSAO + checked blur in front of hard / PCF lighting; SSAA resolve
@pytest.mark.parametrize("scene,shadow,wrap,filt", [("scenes/c1.scene", "hard", 0, 0), ("scenes/c1.scene", "pcf", 0, 0),
                                                      ("scenes/c1_ssao.scene", "hard", 0, 0), ("scenes/c1_ssao.scene", "pcf", 0, 0),
                                                      ("scenes/c4_catbox.scene", "hard", 1, 1)])
def test_replayed_frame_is_the_eager_frame(scene, shadow, wrap, filt, gpu_host, golden):
    host = gpu_host
    sc = host.load_scene(os.path.join(P.REPO, scene), P.ASSETS, wrap, filt)
    try:
        host.render(sc, shadow, True)
        eager = read_all(host, sc)
        cfg = {("scenes/c1.scene", "hard"): "c1_hard", ("scenes/c1.scene", "pcf"): "c1_pcf", ("scenes/c4_catbox.scene", "hard"): "c4_catbox_linear"}.get((scene, shadow))
        if cfg:  # the eager frame is the reference's frame
            for name in EXACT_PLANES:
                if name in golden[cfg]["planes"] and name in eager:
                    assert sha(eager[name]) == golden[cfg]["planes"][name]["sha256"], name
        l0 = host.fgl.launch_count()
        flags = [host.render_replay(sc, shadow, True) for _ in range(4)]
        # the first call may be eager (a cold facade) or already a recording (the facade was warm from an earlier test)
        assert flags[1:] == [True, True, True], (flags, host.replay_fallback_reason())
        assert host.fgl.launch_count() > l0  # replays count the kernels they run
        assert_same(eager, read_all(host, sc), "%s / %s" % (scene, shadow))
        # a read in between must not disturb the next replay, and an eager frame in between must not disturb the recording
        host.render(sc, shadow, True)
        assert host.render_replay(sc, shadow, True) is True
        assert_same(eager, read_all(host, sc), "%s / %s after an eager frame" % (scene, shadow))
    finally:
        sc.free()


def test_moved_camera_and_light_patch_the_recording(gpu_host):
    host = gpu_host
    sc = host.load_scene(os.path.join(P.REPO, "scenes/c1_ssao.scene"), P.ASSETS, 0, 0)
    try:
        poses = [((-1.0, 1.0, 1.0), (0.0, 0.0, -1.0), (2.0, 5.0, 5.0)), ((1.2, 0.6, 0.8), (0.1, -0.2, -1.0), (-1.5, 4.0, 3.0)),
                 ((0.3, 1.4, 1.5), (0.0, -0.1, -1.0), (2.0, 5.0, 5.0))]
        host.render(sc, "hard", True)
        for eye, look, light in poses:
            host.set_camera(sc, eye, look)
            host.set_point_light(sc, light, (1.0, 1.0, 1.0))
            host.render(sc, "hard", True)
            eager = read_all(host, sc)
            assert host.render_replay(sc, "hard", True) in (True, False)  # (False only on a cold facade)
            assert host.render_replay(sc, "hard", True) is True, host.replay_fallback_reason()
            assert_same(eager, read_all(host, sc), "camera %s" % (eye,))
        # another shadow filter is another recording (a pass more: the graph is rebuilt, not patched)
        host.render(sc, "pcf", True)
        eager = read_all(host, sc)
        host.render_replay(sc, "pcf", True)
        assert host.render_replay(sc, "pcf", True) is True, host.replay_fallback_reason()
        assert_same(eager, read_all(host, sc), "pcf after hard")
    finally:
        sc.free()


@pytest.mark.parametrize("scene,shadow", [("scenes/c1_ssao.scene", "pcss"), ("scenes/c2.scene", "pcf")])
def test_unrecordable_frames_fall_back_to_eager_rendering(scene, shadow, gpu_host):
    """PCSS (host read-backs of the chain state) and forward mode with a stochastic filter cannot be recorded: frh_render_replay
    renders them eagerly, says why, and the image is the eager image."""
    host = gpu_host
    sc = host.load_scene(os.path.join(P.REPO, scene), P.ASSETS, 0, 0)
    try:
        host.render(sc, shadow, True)
        names = ["depth", "frame", "frame_u8", "ids_camera"]
        eager = {n: host.fgl.read_plane(n).copy() for n in names}
        flags = [host.render_replay(sc, shadow, True) for _ in range(3)]
        assert flags == [False, False, False]
        assert "recorded frame" in host.replay_fallback_reason() or "record" in host.replay_fallback_reason()
        got = {n: host.fgl.read_plane(n).copy() for n in names}
        for n in names:
            assert P.bits_equal(eager[n], got[n]), n
        # and the context is still good for an ordinary frame
        host.render(sc, shadow, True)
        assert P.bits_equal(eager["frame"], host.fgl.read_plane("frame"))
    finally:
        sc.free()


def test_raw_abi_record_replay_and_refusals():
    """The C ABI itself: record a synthetic frame, replay it, compare with the eager frame; reads / uploads / PCSS inside a
    recording are refused with FGL_ERR_UNSUPPORTED and the recording can be aborted without harming the context."""
    f = B.product_fgl(0)
    try:
        s = SyntheticScene(f, quads=24)
        s.render(320, 200, shadow_mode=B.SHADOW_HARD, ssao=True)
        names = ["depth", "shadow", "ao", "frame", "frame_u8", "ids_camera", "ids_light"]
        eager = {n: f.read_plane(n).copy() for n in names}
        f.frame_record_begin()
        s.render(320, 200, shadow_mode=B.SHADOW_HARD, ssao=True)
        with pytest.raises(B.FglError):
            f.read_plane("depth")
        fid = f.frame_record_end()
        assert fid >= 0
        nodes, launches = f.frame_info(fid)
        assert nodes > 5 and launches > 5
        for _ in range(3):
            f.frame_replay(fid)
        got = {n: f.read_plane(n).copy() for n in names}
        for n in names:
            assert P.bits_equal(eager[n], got[n]), n
        # PCSS inside a recording: refused, abort, then the same frame eagerly
        f.frame_record_begin()
        with pytest.raises(B.FglError) as e:
            s.render(320, 200, shadow_mode=B.SHADOW_PCSS, ssao=True)
        assert "record" in str(e.value)  # (whichever refusal comes first: the larger sample table, or the chain's read-backs)
        f.frame_record_abort()
        s.render(320, 200, shadow_mode=B.SHADOW_HARD, ssao=True)
        assert P.bits_equal(eager["frame"], f.read_plane("frame"))
        # the earlier recording is still valid
        f.frame_replay(fid)
        assert P.bits_equal(eager["frame_u8"], f.read_plane("frame_u8"))
        f.frame_release(fid)
        with pytest.raises(B.FglError):
            f.frame_replay(fid)
    finally:
        f.close()

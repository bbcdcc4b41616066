"""frh_render_replay on a back end that cannot record frames (the CPU oracle answers fgl_frame_record_begin with
FGL_ERR_UNSUPPORTED): every call must fall back to eager rendering, say why, and produce the frame frh_render produces —
the host-side state machine of host/capi.cpp (cold -> warm -> record attempt -> unsupported), exercised without a GPU.
The recording path itself is tested on the GPU (tests/test_gpu_replay.py)."""
import os

import parity as P


def test_replay_falls_back_to_eager_on_a_backend_without_recording(oracle_host):
    host = oracle_host
    sc = host.load_scene(os.path.join(P.REPO, "scenes/c1.scene"), P.ASSETS, 0, 0)
    try:
        host.render(sc, "hard", True)
        names = ["depth", "shadow", "frame", "frame_u8", "ids_camera"]
        eager = {n: host.fgl.read_plane(n).copy() for n in names}
        flags = [host.render_replay(sc, "hard", True) for _ in range(3)]
        assert flags == [False, False, False]
        assert "record" in host.replay_fallback_reason()
        for n in names:
            assert P.bits_equal(eager[n], host.fgl.read_plane(n)), n
        # a changed camera gives the frame another (again refused) recording attempt and still the right image
        host.set_camera(sc, (1.2, 0.6, 0.8), (0.1, -0.2, -1.0))
        host.render(sc, "hard", True)
        moved = {n: host.fgl.read_plane(n).copy() for n in names}
        assert not P.bits_equal(moved["frame_u8"], eager["frame_u8"])
        assert [host.render_replay(sc, "hard", True) for _ in range(3)] == [False, False, False]
        for n in names:
            assert P.bits_equal(moved[n], host.fgl.read_plane(n)), n
    finally:
        sc.free()


def test_oracle_refuses_the_recording_entry_points(oracle_fgl):
    import pytest
    from forkerrenderer_b200 import binding as B
    with pytest.raises(B.FglError):
        oracle_fgl.frame_record_begin()
    with pytest.raises(B.FglError):
        oracle_fgl.frame_replay(0)
    oracle_fgl.frame_record_abort()  # harmless without a recording

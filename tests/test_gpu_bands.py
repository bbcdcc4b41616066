"""Sort-first bands on ONE GPU: two contexts render the two halves of a frame one after the other, the second one
starting its PCSS chain from the first one's blocker count — the result must equal the single-context frame bit for
bit (this is the data path bench.py --gpus N runs with one process per GPU and NCCL in between)."""
import numpy as np
import pytest

import parity as P
from conftest import sha
from forkerrenderer_b200 import binding as B
from forkerrenderer_b200 import multigpu as M
from forkerrenderer_b200.synthetic import SyntheticScene

pytestmark = pytest.mark.gpu


class LocalComm:
    def __init__(self):
        self.box = {}

    def send_int(self, v, dst):
        self.box[dst] = v

    def recv_int(self, src):
        return self.box.pop(src + 1)


@pytest.mark.parametrize("mode,ssao,world", [(B.SHADOW_PCSS, True, 2), (B.SHADOW_PCSS, False, 3), (B.SHADOW_PCF, True, 2), (B.SHADOW_HARD, True, 4)])
def test_bands_equal_full_frame(mode, ssao, world, gpu_fgl):
    W, H = 320, 240
    full_scene = SyntheticScene(gpu_fgl, quads=40)
    full_scene.render(W, H, shadow_mode=mode, ssao=ssao)
    full = gpu_fgl.read_plane("frame_u8")
    comm = LocalComm()
    ctxs = [B.product_fgl(0) for _ in range(world)]
    try:
        for rank, f in enumerate(ctxs):
            r = M.SyntheticRenderer(SyntheticScene(f, quads=40), W, H, shadow_mode=mode, ssao=ssao)
            r0, r1 = M.render_frame(r, rank, world, comm)
            got = f.read_plane("frame_u8")
            assert np.array_equal(got[r0:r1], full[r0:r1]), "band %d of %d differs from the full frame" % (rank, world)
    finally:
        for f in ctxs:
            f.close()


@pytest.mark.parametrize("ssao,world", [(True, 2), (False, 4)])
def test_bands_with_device_side_handoff(ssao, world, gpu_fgl):
    """The same, with the chain state travelling through the contexts' mailboxes (fgl_chain_peer_*: a peer store from the
    band above, a device-side wait in front of the chain kernel) instead of the host — two frames, to cover the epochs."""
    W, H = 320, 240
    full_scene = SyntheticScene(gpu_fgl, quads=40)
    full_scene.render(W, H, shadow_mode=B.SHADOW_PCSS, ssao=ssao)
    full = gpu_fgl.read_plane("frame_u8")
    ctxs = [B.product_fgl(0) for _ in range(world)]
    try:
        boxes = [f.chain_peer_mailbox()[0] for f in ctxs]
        for rank, f in enumerate(ctxs):
            f.chain_peer_connect(next_ptr=boxes[rank + 1] if rank + 1 < world else None, wait_prev=rank > 0)
        renderers = [M.SyntheticRenderer(SyntheticScene(f, quads=40), W, H, shadow_mode=B.SHADOW_PCSS, ssao=ssao) for f in ctxs]
        for frame in range(2):
            for rank, (f, r) in enumerate(zip(ctxs, renderers)):
                r0, r1 = M.render_frame(r, rank, world, None, peer=True)
                got = f.read_plane("frame_u8")
                assert np.array_equal(got[r0:r1], full[r0:r1]), "frame %d: band %d of %d differs from the full frame" % (frame, rank, world)
        assert ctxs[-1].get_chain_blockers() == gpu_fgl.get_chain_blockers()
    finally:
        for f in ctxs:
            f.close()


def test_band_planes_read_back_cleared_outside_the_band(gpu_fgl):
    """A band's passes only write their own rows (+ halo); the clear of the other rows is deferred until somebody reads the
    plane — a host read must still see the reference's clear values there, and the band's rows of the full render."""
    W, H, world, rank = 320, 240, 3, 1
    SyntheticScene(gpu_fgl, quads=40).render(W, H, shadow_mode=B.SHADOW_HARD, ssao=True)
    full = {k: gpu_fgl.read_plane(k) for k in ("normal", "albedo", "ao", "depth")}
    f = B.product_fgl(0)
    try:
        r = M.SyntheticRenderer(SyntheticScene(f, quads=40), W, H, shadow_mode=B.SHADOW_HARD, ssao=True)
        r0, r1 = M.render_frame(r, rank, world, None)
        for k, clear in (("normal", 0.0), ("albedo", 0.0), ("ao", 1.0)):
            got = f.read_plane(k)
            assert np.array_equal(got[r0:r1], full[k][r0:r1]), k
            lo, hi = max(0, r0 - 64), min(H, r1 + 4)  # halo rows hold band-local data
            assert np.all(got[:lo] == clear) and np.all(got[hi:] == clear), k
        assert np.array_equal(f.read_plane("depth"), full["depth"])  # SSAO gathers depth from anywhere: resolved everywhere
    finally:
        f.close()


def test_group_of_one_is_the_standalone_frame(gpu_host, golden):
    """The sort-first group code path (fgl_group_export / connect: band-restricted passes, peer stores — here into the context's
    own planes — epoch flags, rank 0's gathered 8-bit frame) with a group of ONE on this GPU: every plane must be the
    stand-alone frame's, bit for bit, and carry the reference's fingerprints.  (tools/group_check.py does the same with 2 - 8
    processes on as many GPUs.)"""
    import os
    cfg = "c1_ssao_pcss"
    scene, shadow, wrap, filt = P.CONFIGS[cfg]
    sc = gpu_host.load_scene(os.path.join(P.REPO, scene), P.ASSETS, wrap, filt)
    names = ["depth", "shadow", "normal", "worldpos", "lightndc", "ao", "frame", "ids_camera"]
    try:
        gpu_host.render(sc, shadow, True)
        f = gpu_host.fgl
        alone = {n: f.read_plane(n).copy() for n in names + ["frame_u8"]}
        member = gpu_host.group_export(sc)
        gpu_host.group_connect(0, 1, [member], same_process=True)
        try:
            for _ in range(2):
                gpu_host.render(sc, shadow, True)
                img = gpu_host.group_read_frame(sc.height, sc.width)
                assert np.array_equal(img, alone["frame_u8"])
                assert np.array_equal(f.read_plane("frame_u8"), alone["frame_u8"])
                for n in names:
                    assert P.bits_equal(f.read_plane(n), alone[n]), n
            with pytest.raises(B.FglError):
                f.set_row_band(0, 10)  # the band follows from the rank
        finally:
            gpu_host.group_disconnect()
        gpu_host.render(sc, shadow, True)  # and back to a stand-alone context
        assert np.array_equal(f.read_plane("frame_u8"), alone["frame_u8"])
        want = golden[cfg]["planes"]
        for n in ("depth", "shadow", "ao", "ids_camera"):
            assert sha(f.read_plane(n)) == want[n]["sha256"], n
    finally:
        sc.free()

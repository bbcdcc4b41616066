"""Frames in flight: two independent instances of the facade in one process (binding.product_host(instance=k): their own
ForkerGL statics, scene and fgl context), each driven by its own host thread on its own stream, rendering concurrently on the
same GPU.  Every frame of either instance must be the frame the instance renders alone — bit for bit — and the planes the
north star wants exact must carry the reference's fingerprints.  (PCSS frames: the persistent chain kernels of the two
contexts are ordered behind one another on the device, stream.cu launch_chain.)"""
import os
import threading

import numpy as np
import pytest

import parity as P
from conftest import EXACT_PLANES, sha
from forkerrenderer_b200 import binding as B

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", ["c1_ssao_pcss", "c1_pcf"])
def test_two_instances_render_concurrently(cfg, golden):
    scene, shadow, wrap, filt = P.CONFIGS[cfg]
    hosts = [B.product_host(0), B.product_host(1)]
    scenes = [h.load_scene(os.path.join(P.REPO, scene), P.ASSETS, wrap, filt) for h in hosts]
    try:
        alone = []
        for h, sc in zip(hosts, scenes):
            h.render(sc, shadow, True)
            alone.append({n: h.fgl.read_plane(n).copy() for n in ("frame_u8", "frame", "depth", "ao", "ids_camera")})
        assert P.bits_equal(alone[0]["frame"], alone[1]["frame"])
        want = golden[cfg]["planes"]
        for name in ("depth", "ao", "ids_camera"):
            if name in want and name in EXACT_PLANES:
                assert sha(alone[1][name]) == want[name]["sha256"], name
        bad, errors = [], []

        def work(k):
            try:
                for it in range(6):
                    hosts[k].render(scenes[k], shadow, True)
                    for n in ("frame", "depth", "ao"):
                        if not P.bits_equal(hosts[k].fgl.read_plane(n), alone[k][n]):
                            bad.append((k, it, n))
            except Exception as e:
                errors.append(e)

        threads = [threading.Thread(target=work, args=(k,)) for k in range(2)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        assert not errors, errors
        assert not bad, bad
    finally:
        for sc in scenes:
            sc.free()

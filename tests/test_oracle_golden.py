"""Pins the CPU oracle (oracle/oracle.cpp behind the product's own host facade) to the UNMODIFIED reference: the
sha256 fingerprints in tests/golden/reference_hashes.json were produced by oracle/_ref/ref_driver (the reference's
own sources, compiled by oracle/Makefile).  Every plane must match bit for bit — including the 8-bit images, i.e. the
whole mt19937 sample stream is replayed exactly."""
import os

import numpy as np
import pytest

import parity as P
from conftest import sha

FAST = ["c1_hard", "c1_pcf", "c1_pcss", "c1_ssao_pcss", "c2_hard", "c2_pcf", "c4_hard", "c4_catbox_linear", "pbr_hard",
        "fwd_pbr_hard", "fwd_pbr_pcf", "fwd_pbr_pcss", "pbr_ssao_pcss", "catbox_mirrored_linear", "catbox_mirrored_nearest", "catbox_clamp_linear",
        "catbox_clamp_nearest", "catbox_repeat_nearest", "catbox_nowrap_linear", "c1_cam2_pcss", "c1_ortho_hard", "c5_golden"]
SLOW = ["c3_pcss_ssao", "c3_pbr_pcss_ssao"]


def check(cfg, oracle_host, golden):
    got = P.render_host(oracle_host, cfg)
    want = golden[cfg]["planes"]
    bad = []
    for name, fp in want.items():
        if name not in got:
            bad.append(name + " (missing)")
            continue
        a = got[name]
        assert list(a.shape) == fp["shape"], (name, a.shape, fp["shape"])
        if sha(a) != fp["sha256"]:
            bad.append(name)
    assert not bad, "oracle differs from the reference on %s: %s" % (cfg, bad)


@pytest.mark.parametrize("cfg", FAST)
def test_oracle_matches_reference(cfg, oracle_host, golden):
    check(cfg, oracle_host, golden)


@pytest.mark.slow
@pytest.mark.parametrize("cfg", SLOW)
def test_oracle_matches_reference_4k(cfg, oracle_host, golden):
    if not os.environ.get("FGL_SLOW"):
        pytest.skip("about two minutes of CPU; set FGL_SLOW=1")
    check(cfg, oracle_host, golden)


def test_golden_was_generated_by_the_reference_when_it_is_here(golden):
    """If the reference binary is present, re-run it on C1 and compare with the committed fingerprints."""
    if not P.have_ref():
        pytest.skip("oracle/_ref/ref_driver not built")
    ref = P.run_reference("c1_hard")
    for name in ("depth", "frame_u8", "ids_camera", "normal"):
        assert sha(ref[name]) == golden["c1_hard"]["planes"][name]["sha256"], name


def test_generated_c5_input_is_the_one_the_reference_rendered(golden):
    """The C5 assets are generated, not committed: the generator must reproduce the OBJ the golden was made from."""
    assert P.generated_input_md5("c5_golden") == golden["c5_golden"]["generated_obj_md5"]


@pytest.mark.parametrize("kind", sorted(P.BUFFER_KINDS))
def test_oracle_buffer_post_processing_matches_reference(kind, oracle_fgl, golden):
    """Buffer1f/3f::SimpleBlurDenoised and ::TwoPassGaussianBlurDenoised (buffer.cpp:35-98, 140-203): the in-place raster-order
    semantics, against fingerprints of the reference's own Buffer classes on the same LCG inputs."""
    for W, H in P.BUFFER_SHAPES:
        a1, a3 = P.buffer_test_inputs(W, H)
        g1, g3 = P.blur_through_abi(oracle_fgl, P.BUFFER_KINDS[kind], a1, a3)
        want = golden["buffer_ops"]["%s_%dx%d" % (kind, W, H)]
        assert sha(g1) == want["buffer1f"]["sha256"], (kind, W, H, "Buffer1f")
        assert sha(g3) == want["buffer3f"]["sha256"], (kind, W, H, "Buffer3f")

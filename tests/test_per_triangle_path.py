"""The reference's own submission path — Mesh::Draw as a per-face loop of shader.Use / 3 x Shader::ProcessVertex /
ForkerGL::DrawTriangle (reference mesh.cpp:10-25, forkergl.h:74, shader.h:28) — through the facade: the vertex programs run on
the HOST (host/programs.cpp), the triangles reach the back end in batches (fgl_draw_triangles).  The frame must be the one
the reference produced, bit for bit on every plane, exactly as with the indexed fast path.

CPU: the facade over the oracle (pins the host vertex programs + the batching to the reference's fingerprints).
GPU: the facade over the CUDA library (C1 hard + forward PBR with PCF, through k_setup's pre-transformed branch)."""
import numpy as np
import pytest

import parity as P
from conftest import COLOUR_MAX_LSB, EXACT_PLANES, FRAME_F32_MAX_ABS, sha


def render_per_triangle(host, cfg):
    host.set_per_triangle_submission(True)
    try:
        return P.render_host(host, cfg)
    finally:
        host.set_per_triangle_submission(False)


@pytest.mark.parametrize("cfg", ["c1_hard", "fwd_pbr_pcf", "catbox_mirrored_linear"])
def test_oracle_backed_facade_per_triangle(cfg, oracle_host, golden):
    got = render_per_triangle(oracle_host, cfg)
    bad = [name for name, fp in golden[cfg]["planes"].items() if name not in got or sha(got[name]) != fp["sha256"]]
    assert not bad, "per-triangle submission differs from the reference on %s: %s" % (cfg, bad)


@pytest.mark.gpu
@pytest.mark.parametrize("cfg", ["c1_hard", "fwd_pbr_pcf", "c1_pcss"])
def test_cuda_per_triangle(cfg, gpu_host, golden):
    got = render_per_triangle(gpu_host, cfg)
    assert gpu_host.fgl.launch_count() > 0
    want = golden[cfg]["planes"]
    for name in EXACT_PLANES:
        if name in want:
            assert sha(got[name]) == want[name]["sha256"], "%s: %s differs from the reference" % (cfg, name)
    fast = P.render_host(gpu_host, cfg)  # the indexed fast path on the same device: identical images
    for name in ("frame", "frame_u8"):
        assert P.bits_equal(got[name], fast[name]), name
    if P.have_ref():
        ref = P.run_reference(cfg)
        assert P.diff_stats(got["frame"], ref["frame"])["max_abs"] <= FRAME_F32_MAX_ABS
        assert np.abs(got["frame_u8"].astype(int) - ref["frame_u8"].astype(int)).max() <= COLOUR_MAX_LSB

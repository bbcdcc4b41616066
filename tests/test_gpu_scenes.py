"""Whole-frame parity of the CUDA path on the BASELINE configurations, through the host facade (scene / OBJ / TGA
loaders + Render::*), against (a) the committed fingerprints of the UNMODIFIED reference and (b) the reference itself
re-run on this machine when oracle/_ref/ref_driver travelled with the repo (it does under gpurun)."""
import numpy as np
import pytest

import parity as P
from conftest import COLOUR_MAX_LSB, COLOUR_MIN_FRACTION_WITHIN_1LSB, EXACT_PLANES, FRAME_F32_MAX_ABS, sha

pytestmark = pytest.mark.gpu

CONFIGS = ["c1_hard", "c1_pcf", "c1_pcss", "c1_ssao_pcss", "c2_hard", "c2_pcf", "c4_hard", "c4_catbox_linear", "pbr_hard", "c3_pcss_ssao",
           # round 2: forward PBR program (hard / PCF / PCSS), PBR + PCSS + SSAO (1280x800 and 4K), wrap modes 2 / 3, second camera, ortho
           "fwd_pbr_hard", "fwd_pbr_pcf", "fwd_pbr_pcss", "pbr_ssao_pcss", "c3_pbr_pcss_ssao", "catbox_mirrored_linear", "catbox_mirrored_nearest",
           "catbox_clamp_linear", "catbox_clamp_nearest", "catbox_repeat_nearest", "catbox_nowrap_linear", "c1_cam2_pcss", "c1_ortho_hard",
           # C5 (generated OBJ / MTL / TGA / .scene through the facade's loaders; 500 000 triangles: k_raster_small + device-wide scan)
           "c5_golden"]


@pytest.mark.parametrize("cfg", CONFIGS)
def test_scene_matches_reference(cfg, gpu_host, golden):
    got = P.render_host(gpu_host, cfg)
    want = golden[cfg]["planes"]
    # (a) bit-exact planes against the reference's committed fingerprints
    for name in EXACT_PLANES:
        if name in want:
            assert name in got, name
            assert sha(got[name]) == want[name]["sha256"], "%s: %s differs from the reference" % (cfg, name)
    # (b) colour against the reference's actual images
    if not P.have_ref():
        pytest.skip("oracle/_ref/ref_driver is not here: colour planes cannot be compared (bit-exact planes passed)")
    ref = P.run_reference(cfg)
    st = P.diff_stats(got["frame"], ref["frame"])
    assert st["max_abs"] <= FRAME_F32_MAX_ABS, st
    for name in ("frame_u8", "ssaa_u8"):
        if name in ref:
            d = np.abs(got[name].astype(int) - ref[name].astype(int))
            assert d.max() <= COLOUR_MAX_LSB, (cfg, name, int(d.max()))
            assert 1.0 - P.pixel_frac_gt1(got[name], ref[name]) >= COLOUR_MIN_FRACTION_WITHIN_1LSB
            assert (d > 0).mean() < 1e-5, (cfg, name, float((d > 0).mean()))

"""CUDA path vs CPU oracle on the same synthetic submission through the C ABI (both libraries receive identical
calls).  Integer / coverage / depth / G-buffer planes must be bit-exact; the lit frame goes through powf."""
import numpy as np
import pytest

import parity as P
from conftest import COLOUR_MAX_LSB, COLOUR_MIN_FRACTION_WITHIN_1LSB, EXACT_PLANES, FRAME_F32_MAX_ABS
from forkerrenderer_b200 import binding as B
from forkerrenderer_b200.synthetic import SyntheticScene

pytestmark = pytest.mark.gpu

CASES = {
    "hard": dict(shadow_mode=B.SHADOW_HARD),
    "hard_pbr_linear": dict(shadow_mode=B.SHADOW_HARD, pbr=True, filt=B.FILTER_LINEAR),
    "hard_mirrored": dict(shadow_mode=B.SHADOW_HARD, wrap=B.WRAP_MIRRORED_REPEAT, filt=B.FILTER_LINEAR),
    "hard_clamp": dict(shadow_mode=B.SHADOW_HARD, wrap=B.WRAP_CLAMP_TO_EDGE),
    "hard_nowrap": dict(shadow_mode=B.SHADOW_HARD, wrap=B.WRAP_NOWRAP, filt=B.FILTER_LINEAR),
    "pcf": dict(shadow_mode=B.SHADOW_PCF),
    "pcss": dict(shadow_mode=B.SHADOW_PCSS),
    "pcss_ssao": dict(shadow_mode=B.SHADOW_PCSS, ssao=True),
    "pcf_ssao_pbr": dict(shadow_mode=B.SHADOW_PCF, ssao=True, pbr=True),
    "ssaa2": dict(shadow_mode=B.SHADOW_HARD, ssaa=2),
    "ssaa3_odd": dict(shadow_mode=B.SHADOW_HARD, ssaa=3, size=(101, 67)),
    "no_shadow": dict(shadow_mode=B.SHADOW_HARD, shadow=False),
    "forward_hard": dict(shadow_mode=B.SHADOW_HARD, forward=True),
    "forward_pcf": dict(shadow_mode=B.SHADOW_PCF, forward=True),
    "forward_pcss": dict(shadow_mode=B.SHADOW_PCSS, forward=True, size=(300, 180)),
    "forward_pcf_dense": dict(shadow_mode=B.SHADOW_PCF, forward=True, quads=96, size=(200, 120)),
    "odd_size_pcss": dict(shadow_mode=B.SHADOW_PCSS, ssao=True, size=(333, 211)),
    "dense_mesh": dict(shadow_mode=B.SHADOW_PCSS, quads=160, size=(256, 160)),
    # the raster path of the high-triangle-count workload (C5): >= 65 536 triangles per pass switches the small triangles to the
    # thread-per-triangle kernel (k_raster_small), >= 200 000 switches the block-count scan to the device-wide one
    "mesh_80k_pcss_ssao": dict(shadow_mode=B.SHADOW_PCSS, ssao=True, pbr=True, quads=200, size=(256, 160)),
    "mesh_80k_forward_pcf": dict(shadow_mode=B.SHADOW_PCF, forward=True, quads=200, size=(200, 120)),
    "mesh_231k_pcss_ssao": dict(shadow_mode=B.SHADOW_PCSS, ssao=True, pbr=True, quads=340, size=(320, 200)),
    "mesh_231k_forward_pcf": dict(shadow_mode=B.SHADOW_PCF, forward=True, quads=340, size=(200, 120)),
    "mesh_231k_hard_ssaa2": dict(shadow_mode=B.SHADOW_HARD, ssaa=2, quads=340, size=(160, 100)),
}


def run(f, kw):
    kw = dict(kw)
    W, H = kw.pop("size", (320, 200))
    skw = {k: kw.pop(k) for k in ("pbr", "filt", "wrap", "quads") if k in kw}
    s = SyntheticScene(f, **skw)
    s.render(W, H, **kw)
    names = ["depth", "frame", "frame_u8", "ids_camera"]
    if kw.get("shadow", True):
        names += ["shadow", "ids_light"]
    if not kw.get("forward"):
        names += ["normal", "worldpos", "albedo", "emissive", "param", "shadingtype", "ao"] + (["lightndc"] if kw.get("shadow", True) else [])
    if kw.get("ssaa", 1) > 1:
        names.append("ssaa_u8")
    return {n: f.read_plane(n) for n in names}


@pytest.mark.parametrize("case", sorted(CASES))
def test_cuda_matches_oracle(case, gpu_fgl, oracle_fgl):
    got, want = run(gpu_fgl, CASES[case]), run(oracle_fgl, CASES[case])
    assert gpu_fgl.launch_count() > 0
    for name in want:
        if name in EXACT_PLANES:
            assert P.bits_equal(got[name], want[name]), "%s: plane %s is not bit-exact: %s" % (case, name, P.diff_stats(got[name], want[name]))
    st = P.diff_stats(got["frame"], want["frame"])
    assert st["max_abs"] <= FRAME_F32_MAX_ABS, st
    for name in ("frame_u8", "ssaa_u8"):
        if name in want:
            d = np.abs(got[name].astype(int) - want[name].astype(int))
            assert d.max() <= COLOUR_MAX_LSB, (name, int(d.max()))
            assert 1.0 - P.pixel_frac_gt1(got[name], want[name]) >= COLOUR_MIN_FRACTION_WITHIN_1LSB
            assert (d > 0).mean() < 1e-4, (name, float((d > 0).mean()))


def test_two_frames_are_identical(gpu_fgl):
    """Every frame restarts the sample stream (Render::Render runs once per process in the reference) and the
    atomicMin depth|id keys make the rasteriser independent of scheduling: frames are bit-reproducible."""
    s = SyntheticScene(gpu_fgl)
    outs = []
    for _ in range(2):
        s.render(320, 200, shadow_mode=B.SHADOW_PCSS, ssao=True)
        outs.append({n: gpu_fgl.read_plane(n) for n in ("frame", "depth", "ids_camera", "ao", "frame_u8")})
    for n in outs[0]:
        assert P.bits_equal(outs[0][n], outs[1][n]), n


def test_fast_path_without_fp32_frame_gives_the_same_image(gpu_fgl):
    s = SyntheticScene(gpu_fgl)
    s.render(320, 200, shadow_mode=B.SHADOW_PCSS)
    a = gpu_fgl.read_plane("frame_u8")
    s.render(320, 200, shadow_mode=B.SHADOW_PCSS, materialize_frame_f32=False)
    assert np.array_equal(a, gpu_fgl.read_plane("frame_u8"))


def test_row_band_renders_exactly_its_rows(gpu_fgl):
    """Sort-first: a context restricted to a row band produces, on those rows, exactly the full frame's pixels."""
    s = SyntheticScene(gpu_fgl)
    s.render(320, 200, shadow_mode=B.SHADOW_HARD)
    full = {n: gpu_fgl.read_plane(n) for n in ("frame_u8", "depth", "normal", "ids_camera")}
    for band in ((0, 64), (64, 136), (136, 200)):
        s.render(320, 200, shadow_mode=B.SHADOW_HARD, band=band)
        for n in full:
            got = gpu_fgl.read_plane(n)
            assert np.array_equal(got[band[0]:band[1]], full[n][band[0]:band[1]]), (n, band)
    gpu_fgl.set_row_band(0, -1)

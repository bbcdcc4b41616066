"""Shared helpers of the parity tests: run the reference / oracle / CUDA path on a scene, compare planes."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from forkerrenderer_b200 import binding as B  # noqa: E402

ASSETS = os.path.join(REPO, "oracle", "_ref", "assets")
REF_DRIVER = os.path.join(REPO, "oracle", "_ref", "ref_driver")
ORACLE_LIB = os.path.join(REPO, "oracle", "liboracle.so")
ORACLE_HOST_LIB = os.path.join(REPO, "oracle", "liboracle_host.so")

F32_PLANES = ["depth", "shadow", "normal", "worldpos", "lightndc", "albedo", "emissive", "param", "shadingtype", "ao", "frame"]

# scene id -> (scene file, shadow mode, wrap, filter)
CONFIGS = {
    "c1_hard": ("scenes/c1.scene", "hard", 0, 0),
    "c1_pcf": ("scenes/c1.scene", "pcf", 0, 0),
    "c1_pcss": ("scenes/c1.scene", "pcss", 0, 0),
    "c1_ssao_pcss": ("scenes/c1_ssao.scene", "pcss", 0, 0),
    "c2_pcf": ("scenes/c2.scene", "pcf", 0, 0),
    "c2_hard": ("scenes/c2.scene", "hard", 0, 0),
    "c3_pcss_ssao": ("scenes/c3.scene", "pcss", 0, 0),
    "c4_hard": ("scenes/c4.scene", "hard", 0, 0),
    "c4_catbox_linear": ("scenes/c4_catbox.scene", "hard", 1, 1),
    "pbr_hard": ("scenes/pbr.scene", "hard", 0, 0),
    # round 2: forward-mode PBR program, PBR + PCSS + SSAO, the remaining wrap modes, a second camera / light, the orthographic camera
    "fwd_pbr_hard": ("scenes/fwd_pbr.scene", "hard", 0, 0),
    "fwd_pbr_pcf": ("scenes/fwd_pbr.scene", "pcf", 0, 0),
    "fwd_pbr_pcss": ("scenes/fwd_pbr.scene", "pcss", 0, 0),
    "pbr_ssao_pcss": ("scenes/pbr_ssao.scene", "pcss", 0, 0),
    "c3_pbr_pcss_ssao": ("scenes/c3_pbr.scene", "pcss", 0, 0),
    "catbox_mirrored_linear": ("scenes/catbox.scene", "hard", 2, 1),
    "catbox_mirrored_nearest": ("scenes/catbox.scene", "hard", 2, 0),
    "catbox_clamp_linear": ("scenes/catbox.scene", "hard", 3, 1),
    "catbox_clamp_nearest": ("scenes/catbox.scene", "hard", 3, 0),
    "catbox_repeat_nearest": ("scenes/catbox.scene", "hard", 1, 0),
    "catbox_nowrap_linear": ("scenes/catbox.scene", "hard", 0, 1),
    "c1_cam2_pcss": ("scenes/c1_cam2.scene", "pcss", 0, 0),
    "c1_ortho_hard": ("scenes/c1_ortho.scene", "hard", 0, 0),
    # C5 at the size the reference itself rendered: generated OBJ / MTL / TGA / .scene (forkerrenderer_b200/c5.py),
    # 500 x 500 quads = 500 000 triangles, deferred PBR, PCSS + SSAO, Repeat + Linear textures
    "c5_golden": ("@c5_golden", "pcss", 1, 1),
}


def resolve_scene(cfg):
    """(absolute scene path, assets directory) of a config; '@name' scenes are generated on demand (forkerrenderer_b200/c5.py)."""
    scene = CONFIGS[cfg][0]
    if scene.startswith("@"):
        from forkerrenderer_b200 import c5
        root, path = c5.ensure(*c5.INSTANCES[scene[1:]])
        return path, root
    return os.path.join(REPO, scene), ASSETS


def generated_input_md5(cfg):
    """md5 of the generated OBJ of an '@' config (pinned in the golden file: the generator must be deterministic)."""
    import hashlib
    from forkerrenderer_b200 import c5
    path, root = resolve_scene(cfg)
    quads = c5.INSTANCES[CONFIGS[cfg][0][1:]][0]
    h = hashlib.md5()
    with open(os.path.join(root, "obj", "c5_%d" % quads, "field.obj"), "rb") as f:
        for chunk in iter(lambda: f.read(1 << 22), b""):
            h.update(chunk)
    return h.hexdigest()

# parameter sets of the host uniform builders (tests/test_host_math.py): translate xyz, rotY degrees, scale, eye xyz, centre xyz, ratio
MATRIX_CASES = [
    (0, -1, -1, 0, 3, -1, 1, 1, 0, 0, -1, 1.6),
    (0.05, 0, -1, -10, 1, -1, 1, 1, 0, 0, -1, 1.6),
    (0.9, 0, -1, -90, 1, 2, 5, 5, 0, 0, 0, 1.7777778),
    (-0.7, 0, -1, 40, 1, 1.2, 0.6, 0.8, 0.1, -0.2, -1, 1.7777778),
    (0, 0.3, -1, 240, 1, -1.5, 4, 3, 0, 0, 0, 1.0),
    (-1.2, -0.7, -0.6, 30, 0.3, 0.3, 2.5, -4, 0.5, 0.25, 1, 0.5625),
    (0, -0.5, -1, 180, 2.5, 7, 0.01, 0.02, -3, 1, 2, 2.3333333),
    (3.25, -2.5, 1.125, 359.5, 0.015625, -0.001, 12, 0.003, 0, 0, -1, 1.3333334),
]


def run_reference_matrices(case):
    """105 float32 words of the reference's MakeModelMatrix / MakeNormalMatrix / MakeLookAtMatrix / MakePerspectiveMatrix /
    MakeOrthographicMatrix and two products (ref_driver --matrices), as uint32."""
    r = subprocess.run([REF_DRIVER, "--matrices"] + [repr(float(np.float32(v))) for v in case], check=True, stdout=subprocess.PIPE, text=True)
    return np.array([int(w, 16) for w in r.stdout.split()], dtype=np.uint32)


# known-answer vectors of the vertex + fragment programs (ref_driver --fragments / frh_test_fragments): scene, shadow, wrap, filter
FRAGMENT_CASES = {
    "pbr_ssao_pcss": ("scenes/pbr_ssao.scene", "pcss", 0, 0),
    "c2_pcf": ("scenes/c2.scene", "pcf", 0, 0),
    "c3_pbr_hard_repeat_linear": ("scenes/c3_pbr.scene", "hard", 1, 1),
    "catbox_mirrored_linear": ("scenes/catbox.scene", "pcss", 2, 1),
}


def run_reference_fragments(case):
    scene, shadow, wrap, filt = FRAGMENT_CASES[case]
    r = subprocess.run([REF_DRIVER, "--assets", ASSETS, "--scene", os.path.join(REPO, scene), "--out", tempfile.gettempdir(), "--shadow", shadow,
                        "--wrap", str(wrap), "--filter", str(filt), "--quiet", "--fragments"], check=True, stdout=subprocess.PIPE, text=True)
    return np.array([int(w, 16) for w in r.stdout.split()], dtype=np.uint32)


TGA_FILES = ["framebuffer.tga", "framebuffer_SSAA.tga", "shadowmap.tga", "zbuffer.tga", "gbuffer_normal.tga", "gbuffer_worldpos.tga",
             "gbuffer_albedo.tga", "gbuffer_param.tga", "gbuffer_shading_type.tga", "gbuffer_ambient_occlusion.tga"]


def run_reference_tga(cfg):
    """{file name: md5} of the TGA files the reference's own Output::* writes for a config (ref_driver --tga)."""
    import hashlib
    scene, shadow, wrap, filt = CONFIGS[cfg]
    out_dir = os.path.join(tempfile.gettempdir(), "fgl_ref_tga_" + cfg)
    os.makedirs(out_dir, exist_ok=True)
    tga_dir = os.path.join(ASSETS, "output")
    for f in TGA_FILES:
        if os.path.exists(os.path.join(tga_dir, f)):
            os.unlink(os.path.join(tga_dir, f))
    subprocess.run([REF_DRIVER, "--assets", ASSETS, "--scene", os.path.join(REPO, scene), "--out", out_dir, "--shadow", shadow, "--wrap", str(wrap),
                    "--filter", str(filt), "--quiet", "--tga"], check=True, stdout=subprocess.DEVNULL)
    return {f: hashlib.md5(open(os.path.join(tga_dir, f), "rb").read()).hexdigest() for f in TGA_FILES if os.path.exists(os.path.join(tga_dir, f))}


BUFFER_SHAPES = [(5, 4), (33, 7), (1, 9), (9, 1), (64, 48)]
BUFFER_KINDS = {"simple": B.BLUR_SIMPLE_3X3, "gauss": B.BLUR_TWO_PASS_GAUSSIAN}


def buffer_test_inputs(W, H):
    """The deterministic Buffer1f / Buffer3f contents of `ref_driver --buffer-test` (an LCG): (H, W) and (H, W, 3) float32."""
    s = np.uint32(12345 + W * 131 + H)
    out = np.empty(W * H * 4, dtype=np.float32)
    a, c = np.uint32(1664525), np.uint32(1013904223)
    with np.errstate(over="ignore"):
        for i in range(out.size):
            s = np.uint32(s * a + c)
            out[i] = np.float32((int(s) >> 8) & 0xffffff) / np.float32(16777216.0)
    return out[: W * H].reshape(H, W).copy(), out[W * H:].reshape(H, W, 3).copy()


def run_reference_buffer_ops(out_dir=None):
    """{"simple_WxH" / "gauss_WxH": (Buffer1f result, Buffer3f result)} from the UNMODIFIED reference's Buffer classes."""
    out_dir = out_dir or os.path.join(tempfile.gettempdir(), "fgl_ref_buffer_ops")
    os.makedirs(out_dir, exist_ok=True)
    subprocess.run([REF_DRIVER, "--buffer-test", out_dir], check=True)
    res = {}
    for kind in BUFFER_KINDS:
        for W, H in BUFFER_SHAPES:
            raw = np.fromfile(os.path.join(out_dir, "buffer_%s_%dx%d.raw" % (kind, W, H)), dtype=np.float32)
            res["%s_%dx%d" % (kind, W, H)] = (raw[: W * H].reshape(H, W), raw[W * H:].reshape(H, W, 3))
    return res


def blur_through_abi(fgl, kind, a1, a3):
    """Buffer1f / Buffer3f post-processing through the C ABI: the AO plane and the albedo plane stand in for the buffers."""
    H, W = a1.shape
    fgl.init_geometry_buffers(W, H)
    fgl.write_plane("ao", a1)
    fgl.write_plane("albedo", a3)
    fgl.blur(B.PLANE_AO, kind)
    fgl.blur(B.PLANE_ALBEDO, kind)
    return fgl.read_plane("ao"), fgl.read_plane("albedo")


def have_ref():
    return os.path.exists(REF_DRIVER) and os.path.isdir(os.path.join(ASSETS, "obj"))


def run_reference(cfg, out_dir=None, ids=True):
    """Runs the UNMODIFIED reference (oracle/_ref/ref_driver) and returns {plane: array} + meta."""
    scene, shadow, wrap, filt = CONFIGS[cfg]
    out_dir = out_dir or os.path.join(tempfile.gettempdir(), "fgl_ref_" + cfg)
    os.makedirs(out_dir, exist_ok=True)
    meta_path = os.path.join(out_dir, "meta.json")
    if not os.path.exists(meta_path):
        scene_path, assets = resolve_scene(cfg)
        cmd = [REF_DRIVER, "--assets", assets, "--scene", scene_path, "--out", out_dir, "--shadow", shadow,
               "--wrap", str(wrap), "--filter", str(filt), "--quiet"] + (["--ids"] if ids else [])
        subprocess.run(cmd, check=True, stdout=subprocess.DEVNULL)
    meta = json.load(open(meta_path))
    W, H = meta["width"], meta["height"]
    out = {"meta": meta}
    for name in F32_PLANES:
        p = os.path.join(out_dir, name + ".f32")
        if os.path.exists(p):
            a = np.fromfile(p, dtype=np.float32)
            out[name] = a.reshape(H, W, 3) if a.size == W * H * 3 else a.reshape(H, W)
    out["frame_u8"] = np.fromfile(os.path.join(out_dir, "frame.u8"), dtype=np.uint8).reshape(H, W, 3)
    p = os.path.join(out_dir, "ssaa.u8")
    if os.path.exists(p):
        out["ssaa_u8"] = np.fromfile(p, dtype=np.uint8).reshape(meta["out_height"], meta["out_width"], 3)
    for name in ("ids_camera", "ids_light"):
        p = os.path.join(out_dir, name + ".i32")
        if os.path.exists(p):
            out[name] = np.fromfile(p, dtype=np.int32).reshape(H, W)
    return out


def render_host(host, cfg, planes=None, materialize=True):
    """Renders a config through a Host facade (product or oracle-linked) and reads the planes back."""
    scene, shadow, wrap, filt = CONFIGS[cfg]
    scene_path, assets = resolve_scene(cfg)
    sc = host.load_scene(scene_path, assets, wrap, filt)
    try:
        host.render(sc, shadow, materialize)
        f = host.fgl
        names = planes or (["depth", "shadow", "frame", "frame_u8", "ids_camera", "ids_light"] +
                           (["normal", "worldpos", "lightndc", "albedo", "emissive", "param", "shadingtype", "ao"] if sc.deferred else []) +
                           (["ssaa_u8"] if sc.ssaa else []))
        out = {n: f.read_plane(n) for n in names}
        out["scene"] = dict(width=sc.width, height=sc.height, triangles=sc.triangles, deferred=sc.deferred)
        return out
    finally:
        sc.free()


def bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    if a.shape != b.shape or a.dtype != b.dtype:
        return False
    return bool(np.array_equal(a.view(np.uint8), b.view(np.uint8)))


def diff_stats(a, b):
    """dict with max abs diff, count of differing elements, fraction; ints compared exactly."""
    if a.shape != b.shape:
        return dict(shape_mismatch=(a.shape, b.shape))
    if a.dtype == np.float32:
        neq = a.view(np.uint32) != b.view(np.uint32)
        with np.errstate(invalid="ignore", over="ignore"):
            d = np.abs(a.astype(np.float64) - b.astype(np.float64))
        d = np.where(np.isfinite(d), d, 0)
        return dict(n_diff=int(neq.sum()), frac=float(neq.mean()), max_abs=float(d.max()))
    d = np.abs(a.astype(np.int64) - b.astype(np.int64))
    return dict(n_diff=int((d > 0).sum()), frac=float((d > 0).mean()), max_abs=int(d.max()),
                frac_gt1=float((d > 1).mean()))


def pixel_frac_gt1(a, b):
    """fraction of PIXELS with any channel differing by more than 1 LSB (the north-star colour criterion)."""
    d = np.abs(a.astype(np.int64) - b.astype(np.int64))
    if d.ndim == 3:
        d = d.max(axis=2)
    return float((d > 1).mean())

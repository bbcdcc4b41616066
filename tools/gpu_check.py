"""Development report: CUDA path vs oracle / reference with per-plane statistics.  python tools/gpu_check.py [cfg ...]"""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import parity as P
from forkerrenderer_b200 import binding as B
from forkerrenderer_b200.synthetic import SyntheticScene

def report(tag, got, want):
    ok = True
    for k in want:
        if k in ("meta", "scene") or k not in got:
            continue
        st = P.diff_stats(got[k], want[k])
        extra = ""
        if k.endswith("_u8") and "shape_mismatch" not in st:
            extra = " px>1LSB=%.5f%%" % (100 * P.pixel_frac_gt1(got[k], want[k]))
        flag = "OK " if st.get("n_diff", 1) == 0 else "DIFF"
        print("  [%s] %-12s %s %s%s" % (tag, k, flag, st, extra), flush=True)
        ok &= st.get("n_diff", 1) == 0
    return ok

def synthetic():
    gpu = B.product_fgl(0)
    orc = B.Fgl(P.ORACLE_LIB)
    names = ["depth", "shadow", "normal", "worldpos", "lightndc", "albedo", "emissive", "param", "shadingtype", "ao", "frame", "frame_u8", "ids_camera", "ids_light"]
    for kw in (dict(shadow_mode=B.SHADOW_HARD), dict(shadow_mode=B.SHADOW_HARD, pbr=True, filt=B.FILTER_LINEAR),
               dict(shadow_mode=B.SHADOW_PCF), dict(shadow_mode=B.SHADOW_PCSS), dict(shadow_mode=B.SHADOW_PCSS, ssao=True),
               dict(shadow_mode=B.SHADOW_HARD, ssaa=2), dict(shadow_mode=B.SHADOW_HARD, forward=True)):
        print("synthetic", kw, flush=True)
        outs = []
        for f in (gpu, orc):
            skw = {k: kw[k] for k in ("pbr", "filt") if k in kw}
            rkw = {k: v for k, v in kw.items() if k not in skw}
            t = time.time()
            s = SyntheticScene(f, **skw)
            s.render(320, 200, **rkw)
            nm = [n for n in names if not (kw.get("forward") and n in ("normal", "worldpos", "lightndc", "albedo", "emissive", "param", "shadingtype", "ao"))]
            if kw.get("ssaa"): nm = nm + ["ssaa_u8"]
            outs.append({n: f.read_plane(n) for n in nm})
            print("   %s: %.3fs" % (f.backend, time.time() - t), flush=True)
        report("syn", outs[0], outs[1])

def scenes(cfgs):
    host = B.product_host()
    for cfg in cfgs:
        print("scene", cfg, flush=True)
        t = time.time(); ref = P.run_reference(cfg); t_ref = time.time() - t
        t = time.time(); got = P.render_host(host, cfg); t_gpu = time.time() - t
        print("   reference %.2fs (frame %.3fs), cuda path incl. load+readback %.2fs" % (t_ref, ref["meta"]["t_frame"], t_gpu), flush=True)
        report(cfg, got, ref)

if __name__ == "__main__":
    args = sys.argv[1:]
    if not args or "synthetic" in args:
        synthetic()
    cfgs = [a for a in args if a in P.CONFIGS]
    if cfgs:
        scenes(cfgs)

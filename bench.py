#!/usr/bin/env python
"""bench.py — frames/s and Mpixels/s of the raster-and-shade frame (Render::Render, reference render.cpp:40-58).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c1_ssao|c4|c5] [--impl ours|reference]

One step = one frame of the workload scene: shadow pass, geometry pass, SSAO + in-place Gaussian, deferred lighting
with the scene's shadow filter, 8-bit quantise (+ SSAA resolve).  Geometry and textures are resident on the device
(they are the renderer's "model"; the reference's own stopwatch lines exclude scene loading too).

  value   whole-frame throughput in Mpixels/s (output pixels), device-timed with CUDA events, inputs resident.
  e2e     the same metric through the reference-facing facade call (frh_render = Render::Preconfigure + Render::Render)
          including, every step, the host->device copy of the frame's draw commands / uniforms and the device->host
          read of the finished 8-bit frame into host memory; wall-clock timed around the call + read.
  roofline   dominant kernel (per-kernel CUDA events from the library's own instrumentation) against the measured HBM
             copy bandwidth of MEASURED_PEAKS.json.
  cpu_baseline   the UNMODIFIED reference (oracle/_ref/ref_driver, single-threaded by construction) on a bounded
                 sample: the same scene at a reduced resolution, compared in Mpixels/s.

--impl reference times only the CPU reference (rank 0), same metric/unit/config.
"""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (scene file, shadow mode, wrap, filter, description)
    "c1": ("scenes/c1.scene", "hard", 0, 0, "scenes/c1.scene: Mary + plane, 1280x800 deferred, hard shadow"),
    "c1_ssao": ("scenes/c1_ssao.scene", "pcss", 0, 0, "scenes/c1_ssao.scene: Mary + plane, 1280x800 deferred, PCSS + SSAO"),
    "c3": ("scenes/c3.scene", "pcss", 0, 0, "scenes/c3.scene: plane + Mary + diablo_pose + great_sword, 3840x2160 deferred, PCSS + SSAO (two-pass Gaussian)"),
    "c3_pbr": ("scenes/c3_pbr.scene", "pcss", 0, 0, "scenes/c3_pbr.scene: C3 + chalkboard (Cook-Torrance), 3840x2160 deferred, PCSS + SSAO"),
    "c2": ("scenes/c2.scene", "pcf", 0, 0, "scenes/c2.scene: plane + african_head, 1920x1080 forward Blinn-Phong + normal/specular maps, PCF"),
    "c4": ("scenes/c4_catbox.scene", "hard", 1, 1, "scenes/c4_catbox.scene: SSAA 2x (2560x1600 raster), Repeat + Linear textures"),
}
ASSETS = os.path.join(REPO, "oracle", "_ref", "assets")
REF_DRIVER = os.path.join(REPO, "oracle", "_ref", "ref_driver")


def measured_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.proc = gpu_index, [], False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def reduced_scene(scene_path, factor):
    """Same scene at 1/factor of the linear resolution (bounded CPU sample).  Returns (path, width, height)."""
    txt = open(os.path.join(REPO, scene_path)).read()
    m = re.search(r"^screen\s+(\d+)\s+(\d+)", txt, re.M)
    w, h = int(m.group(1)) // factor, int(m.group(2)) // factor
    txt = re.sub(r"^screen\s+\d+\s+\d+", "screen %d %d" % (w, h), txt, flags=re.M)
    fd, path = tempfile.mkstemp(suffix=".scene", prefix="fgl_bench_")
    os.write(fd, txt.encode())
    os.close(fd)
    return path, w, h


def time_reference(workload, frames, factor):
    """Runs the unmodified reference on the (reduced) scene; returns per-frame seconds and the pixel count."""
    scene, shadow, wrap, filt, _ = WORKLOADS[workload]
    if not os.path.exists(REF_DRIVER):
        raise RuntimeError("oracle/_ref/ref_driver is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists")
    path, w, h = reduced_scene(scene, factor)
    out = tempfile.mkdtemp(prefix="fgl_bench_ref_")
    r = subprocess.run([REF_DRIVER, "--assets", ASSETS, "--scene", path, "--out", out, "--shadow", shadow, "--wrap", str(wrap),
                        "--filter", str(filt), "--frames", str(frames), "--quiet"], check=True, stdout=subprocess.PIPE, text=True)
    os.unlink(path)
    times = [json.loads(l)["t_frame"] for l in r.stdout.splitlines() if l.startswith("{")]
    return times, w, h


def cpu_factor(workload):
    return {"c3": 4, "c3_pbr": 4, "c1_ssao": 2}.get(workload, 1)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.workload in SYNTH:
        ts = [time_oracle_port(args.workload) for _ in range(args.steps)]
        w, h = ts[0][1], ts[0][2]
        ms = 1e3 * sum(t[0] for t in ts) / len(ts)
        mpx = w * h / 1e6 / (ms / 1e3)
        sample = "oracle port on the synthetic scene reduced to %d triangles at %dx%d, %d frames, single thread" % (ts[0][3], w, h, len(ts))
        kind, desc = "port", SYNTH[args.workload][4]
    else:
        factor = cpu_factor(args.workload)
        times, w, h = time_reference(args.workload, args.warmup + args.steps, factor)
        t = times[args.warmup:]
        ms = 1e3 * sum(t) / len(t)
        mpx = w * h / 1e6 / (ms / 1e3)
        sample = "%s at %dx%d (1/%d linear resolution), %d frames, single thread" % (WORKLOADS[args.workload][0], w, h, factor, len(t))
        kind, desc = "reference", WORKLOADS[args.workload][4]
    line = {"impl": "reference", "metric": "mpixels_per_s", "value": mpx, "unit": "Mpixels/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "reference assets (obj/*), synthetic camera/light of the scene file",
            "config": {"workload": desc, "sample": sample},
            "frames_per_s": 1e3 / ms,
            "cpu_baseline": {"value": mpx, "unit": "Mpixels/s", "cores": 1, "kind": kind, "sample": sample},
            "e2e": {"value": mpx, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


SYNTH = {
    # name: (quads per side, width, height, render kwargs, description, cpu sample (quads, width, height))
    "c5": (2237, 7680, 4320, dict(shadow_mode=2, ssao=True), "synthetic 10.0 M-triangle height field + plane, deferred PBR, PCSS + SSAO, 7680x4320",
           (280, 960, 540)),
    "c5_small": (700, 1920, 1080, dict(shadow_mode=2, ssao=True), "synthetic 0.98 M-triangle height field + plane, deferred PBR, PCSS + SSAO, 1920x1080",
                 (175, 480, 270)),
}


def make_renderer(args, local):
    from forkerrenderer_b200 import binding as B
    from forkerrenderer_b200 import multigpu as M
    if args.workload in SYNTH:
        from forkerrenderer_b200.synthetic import SyntheticScene
        quads, W, H, kw, desc, _ = SYNTH[args.workload]
        fgl = B.product_fgl(local)
        scene = SyntheticScene(fgl, quads=quads, pbr=True, tex_size=256)
        r = M.SyntheticRenderer(scene, W, H, materialize_frame_f32=False, **kw)
        info = dict(desc=desc, triangles=scene.triangles, out_w=W, out_h=H, ssaa=False, free=lambda: None)
        return r, info
    scene_file, shadow, wrap, filt, desc = WORKLOADS[args.workload]
    host = B.product_host()
    sc = host.load_scene(os.path.join(REPO, scene_file), ASSETS, wrap, filt)
    r = M.FacadeRenderer(host, sc, shadow, materialize=False)
    info = dict(desc=desc, triangles=sc.triangles, out_w=sc.width, out_h=sc.height, ssaa=bool(sc.ssaa), free=sc.free)
    return r, info


def time_oracle_port(workload):
    """C5 has no .scene file the reference could load: its CPU baseline is the oracle port on a reduced sample."""
    from forkerrenderer_b200 import binding as B
    from forkerrenderer_b200.synthetic import SyntheticScene
    quads, W, H = SYNTH[workload][5]
    orc = B.Fgl(os.path.join(REPO, "oracle", "liboracle.so"))
    scene = SyntheticScene(orc, quads=quads, pbr=True, tex_size=256)
    t0 = time.perf_counter()
    scene.render(W, H, **SYNTH[workload][3])
    orc.read_plane("frame_u8")
    dt = time.perf_counter() - t0
    orc.close()
    return dt, W, H, scene.triangles


def _dbg(msg):
    if os.environ.get("FGL_BENCH_DEBUG"):
        print("[bench rank %s] %s" % (os.environ.get("RANK", "0"), msg), file=sys.stderr, flush=True)


def run_ours(args):
    import numpy as np
    import torch
    from forkerrenderer_b200 import binding as B
    from forkerrenderer_b200 import multigpu as M

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL writes its debug output — at NCCL_DEBUG=VERSION / WARN that includes a version banner — to STDOUT unless told
        # otherwise; rank 0's stdout is the one JSON line
        # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    os.environ["FGL_DEVICE"] = str(local)

    _dbg("process group up, building the renderer")
    r, info = make_renderer(args, local)
    fgl = r.fgl
    W, H = r.width, r.height            # raster size (output size x SSAA factor)
    out_px = info["out_w"] * info["out_h"]
    stream = torch.cuda.Stream()
    fgl.set_stream(stream.cuda_stream)
    comm = M.TorchComm(dist, torch.device("cuda", local)) if world > 1 else None
    # PCSS frames: the chain state goes from band to band through peer memory (device-side wait / peer store); --handoff host
    # keeps the NCCL send / recv of one integer per band
    peer = bool(world > 1 and r.pcss and args.handoff == "peer" and M.setup_peer_handoff(fgl, dist, rank, world, H))
    _dbg("chain hand-off: %s" % ("peer memory" if peer else "host"))
    r0, r1, per = M.band_rows(H, world, rank)
    band = torch.empty((per, W, 3), dtype=torch.uint8, device="cuda") if world > 1 else None

    def barrier():
        fgl.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # L2 flush between timed iterations for workloads whose planes could sit in the 126 MB L2
    plane_bytes = W * H * 100 // world
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if plane_bytes < (512 << 20) else None

    def flush_l2():
        if flush_buf is not None:
            flush_buf.fill_(1)
            torch.cuda.synchronize()

    def frame(gather=True):
        """One frame; with several GPUs: this rank's band, the chain hand-off, and the NCCL gather of the 8-bit bands."""
        with torch.cuda.stream(stream):
            M.render_frame(r, rank, world, comm, band_out=(band.data_ptr(), band.numel()) if world > 1 else None, peer=peer)
            if world > 1 and gather:
                return M.gather_bands(dist, torch, band, H, W, world)
        return None

    for _ in range(args.warmup):
        frame()
    barrier()
    _dbg("warm-up done")

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    l0 = fgl.launch_count()
    ev = []
    for _ in range(args.steps):
        flush_l2()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        frame()
        e1.record(stream)
        ev.append((e0, e1))
    barrier()
    launches = fgl.launch_count() - l0
    ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)

    # ---- end to end: facade call(s) + the finished 8-bit frame in host memory, wall clock --------------------------
    e2e_t = []
    d2h = 0
    frame_hash = None
    host_frame = None  # page-locked destination of the gathered frame (rank 0)
    E2E_WARM = 2  # untimed passes: the first one allocates the page-locked frame buffer and fingerprints the frame
    for i in range(args.steps + E2E_WARM):
        flush_l2()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        full = frame()
        if world > 1:
            if rank == 0:
                if host_frame is None:
                    host_frame = torch.empty(full.shape, dtype=torch.uint8, pin_memory=True)
                if os.environ.get("FGL_BENCH_DEBUG"):
                    stream.synchronize()
                    _dbg("e2e step %d: frame + gather %.3f ms" % (i, 1e3 * (time.perf_counter() - t0)))
                with torch.cuda.stream(stream):
                    host_frame.copy_(full, non_blocking=True)
                stream.synchronize()
                img = host_frame
                d2h = img.numel()
            else:
                stream.synchronize()
        else:
            img = fgl.read_plane("ssaa_u8" if info["ssaa"] else "frame_u8", pinned=True)  # page-locked host buffer (fgl_host_alloc)
            d2h = int(img.nbytes)
        t1 = time.perf_counter()
        _dbg("e2e step %d: %.3f ms" % (i, 1e3 * (t1 - t0)))
        if i >= E2E_WARM:
            e2e_t.append(t1 - t0)
        elif i == 0 and rank == 0:  # the untimed first pass: fingerprint of the finished frame (must not depend on the number of GPUs)
            import hashlib
            frame_hash = hashlib.sha256(np.ascontiguousarray(img.numpy() if hasattr(img, "numpy") else img).tobytes()).hexdigest()
    e2e_ms = 1e3 * sum(e2e_t) / len(e2e_t)
    clocks = sampler.finish()
    _dbg("timed loops done: %.3f ms device, %.3f ms e2e" % (ms, e2e_ms))

    # ---- per-kernel breakdown (library instrumentation, separate frames) -----------------------------------------
    barrier()
    fgl.enable_timing(True)
    fgl.reset_timings()
    nprof = 3
    for _ in range(nprof):
        flush_l2()
        if world > 1:
            dist.barrier()  # keep the ranks within one frame of each other (the hand-off mailboxes hold 16 frames)
        frame(gather=False)
    fgl.sync()
    kern = fgl.timings()
    fgl.enable_timing(False)
    fgl.reset_timings()
    for k in kern:
        k["ms_per_frame"] = k["ms_total"] / nprof
    kern.sort(key=lambda k: -k["ms_total"])
    peak, peak_src = measured_peaks()
    byte_kernels = [k for k in kern if k["algorithmic_bytes"] > 0]
    top = byte_kernels[0] if byte_kernels else kern[0]
    t_launch = top["ms_total"] / top["launches"] / 1e3
    bytes_launch = top["algorithmic_bytes"] / top["launches"]
    achieved = bytes_launch / t_launch / 1e9
    total_kernel_ms = max(1e-9, sum(k["ms_per_frame"] for k in kern))
    traffic = None
    try:  # DRAM bytes per launch of that kernel from the committed ncu capture of the same workload (null if there is none)
        traffic = json.load(open(os.path.join(REPO, "profiles", "r01_ncu_traffic.json"))).get(args.workload, {}).get(top["name"])
    except Exception:
        pass
    roofline = {"kernel": top["name"], "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_launch,
                "avg_launch_ms": t_launch * 1e3, "share_of_frame": top["ms_per_frame"] / total_kernel_ms,
                "note": "largest bandwidth-bound kernel; the per-kernel table is in `kernels` (kernels without a byte figure are latency/ALU bound)"}

    if world > 1:
        t = torch.tensor([ms, e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])

    _dbg("reporting")
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            if args.workload in SYNTH:
                dt, cw, ch, ctris = time_oracle_port(args.workload)
                cpu = {"value": cw * ch / 1e6 / dt, "unit": "Mpixels/s", "cores": 1, "kind": "port", "ms_per_frame": 1e3 * dt,
                       "sample": "oracle port (oracle/liboracle.so) on the same synthetic scene reduced to %d triangles at %dx%d, 1 frame, single thread" % (ctris, cw, ch)}
            else:
                factor = cpu_factor(args.workload)
                times, cw, ch = time_reference(args.workload, 1, factor)
                cpu = {"value": cw * ch / 1e6 / times[0], "unit": "Mpixels/s", "cores": 1, "kind": "reference", "ms_per_frame": 1e3 * times[0],
                       "sample": "%s at %dx%d (1/%d linear resolution), 1 frame of oracle/_ref/ref_driver, single thread (the reference has no threads)"
                                 % (WORKLOADS[args.workload][0], cw, ch, factor)}
        line = {"metric": "mpixels_per_s", "value": out_px / 1e6 / (ms / 1e3), "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
                "vs_baseline": None, "dtype": "f32", "data": "reference assets (obj/*) / procedural mesh, camera and light of the scene",
                "config": {"workload": info["desc"], "triangles": info["triangles"], "width": info["out_w"], "height": info["out_h"],
                           "partition": ("sort-first row bands, %d rows per GPU, geometry replicated, RGB8 bands all-gathered with NCCL, PCSS chain state handed on through %s"
                                         % (per, "peer memory (device-side wait)" if peer else "the host (NCCL send/recv)")) if world > 1 else "single GPU",
                           "l2": "explicit 256 MiB flush between timed frames" if flush_buf is not None else "planes (%.0f MB per GPU) exceed the 126 MB L2" % (plane_bytes / 1e6)},
                "frames_per_s": 1e3 / ms,
                "e2e": {"value": out_px / 1e6 / (e2e_ms / 1e3), "unit": "Mpixels/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": 2 * 512 * 8, "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches), "frame_sha256": frame_hash, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "kernels": [{"name": k["name"], "ms_per_frame": round(k["ms_per_frame"], 4), "launches_per_frame": k["launches"] / nprof,
                             "gbps": (k["algorithmic_bytes"] / max(1e-12, k["ms_total"] / 1e3) / 1e9) if k["algorithmic_bytes"] else None}
                            for k in kern]}
        print(json.dumps(line), flush=True)
    info["free"]()
    if world > 1:
        dist.destroy_process_group()
    return 0


def h2d_bytes(sc):
    # per frame the facade uploads one DrawCmd record per mesh per raster pass (shadow + geometry/forward)
    return 0 if sc is None else 2 * 512 * 8


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS) + sorted(SYNTH))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--handoff", default="peer", choices=["peer", "host"], help="N > 1, PCSS: how the chain state travels between the bands")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 0)
    if args.impl == "reference":
        if args.steps > 3:
            args.steps = 3  # each reference frame is seconds of CPU time; keep the arm within minutes
        args.warmup = min(args.warmup, 1)
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

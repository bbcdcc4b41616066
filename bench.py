#!/usr/bin/env python
"""bench.py — frames/s and Mpixels/s of the raster-and-shade frame (Render::Render, reference render.cpp:40-58).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c1|c2|c4|c5|...] [--impl ours|reference]

One step = one frame of the workload scene: shadow pass, geometry pass, SSAO + in-place Gaussian, deferred lighting
with the scene's shadow filter, 8-bit quantise (+ SSAA resolve).  Geometry and textures are resident on the device
(they are the renderer's "model"; the reference's own stopwatch lines exclude scene loading too).
Default workload: C3 (4K PCSS + SSAO, BASELINE.json configs[2]) on one GPU, C5 (8K, 10 M triangles, configs[4] — the
north star's scaling configuration) on several.  Every workload is a .scene file that goes through the facade's loaders
(frh_scene_load); C5's OBJ / MTL / TGA / .scene are generated on the spot (forkerrenderer_b200/c5.py).

  value      whole-frame throughput in Mpixels/s (output pixels), device-timed with CUDA events, inputs resident.
  e2e        the same metric through the reference-facing facade call (frh_render = Render::Preconfigure + Render::Render)
             including, every step, the host->device copy of the frame's draw commands / uniforms and the device->host read
             of the finished 8-bit frame into page-locked host memory; wall-clock timed around the call + read.
  roofline   the kernel with the LARGEST share of the frame (per-kernel CUDA events from the library's own instrumentation,
             taken on serialised frames) against the measured HBM copy bandwidth of MEASURED_PEAKS.json; `limiter` says what
             bounds it when that is not HBM, `counters` quotes the committed ncu capture (profiles/).
  cpu_baseline   the UNMODIFIED reference (oracle/_ref/ref_driver, single-threaded by construction) on a bounded sample.

--impl reference times only the CPU reference (rank 0) on the same config, metric and unit.
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
    # name: (scene file or '@generated instance', shadow mode, wrap, filter, description)
    "c1": ("scenes/c1.scene", "hard", 0, 0, "scenes/c1.scene: Mary + plane, 1280x800 deferred, hard shadow"),
    "c1_ssao": ("scenes/c1_ssao.scene", "pcss", 0, 0, "scenes/c1_ssao.scene: Mary + plane, 1280x800 deferred, PCSS + SSAO"),
    "c3": ("scenes/c3.scene", "pcss", 0, 0, "scenes/c3.scene: plane + Mary + diablo_pose + great_sword, 3840x2160 deferred, PCSS + SSAO (two-pass Gaussian)"),
    "c3_pbr": ("scenes/c3_pbr.scene", "pcss", 0, 0, "scenes/c3_pbr.scene: C3 + chalkboard (Cook-Torrance), 3840x2160 deferred, PCSS + SSAO"),
    "c2": ("scenes/c2.scene", "pcf", 0, 0, "scenes/c2.scene: plane + african_head, 1920x1080 forward Blinn-Phong + normal/specular maps, PCF"),
    "c4": ("scenes/c4_catbox.scene", "hard", 1, 1, "scenes/c4_catbox.scene: SSAA 2x (2560x1600 raster), Repeat + Linear textures"),
    "c5": ("@c5", "pcss", 1, 1, "C5: generated 2237x2237-quad height field OBJ (10 008 338 triangles) + plane, deferred PBR, PCSS + SSAO, Repeat + Linear, 7680x4320"),
    "c5_hard": ("@c5", "hard", 1, 1, "C5 geometry with the hard-shadow filter (no sample-stream chain): generated 10 008 338-triangle OBJ + plane, deferred PBR, hard shadows + SSAO, 7680x4320"),
    "c5_pcf": ("@c5", "pcf", 1, 1, "C5 geometry with the PCF filter (closed-form sample offsets, no chain): generated 10 008 338-triangle OBJ + plane, deferred PBR, PCF + SSAO, 7680x4320"),
    "c5_small": ("@c5_small", "pcss", 1, 1, "C5 reduced: generated 700x700-quad height field OBJ (980 000 triangles) + plane, deferred PBR, PCSS + SSAO, 1920x1080"),
}
ASSETS = os.path.join(REPO, "oracle", "_ref", "assets")
REF_DRIVER = os.path.join(REPO, "oracle", "_ref", "ref_driver")

# What bounds each kernel when it is not HBM bandwidth; quoted next to the roofline figures.
LIMITERS = {
    "pcss_chain": "latency: persistent kernel, one super-chunk per iteration, grid-wide barriers + ordered composition of the segment tables",
    "ssao": "instruction issue: the projection arithmetic of 32 samples per pixel",
    "pcss_visibility": "instruction issue + sample-table stream (warp per pixel, 96 taps)",
    "pcf_visibility": "instruction issue + sample-table stream (warp per site, 64 taps)",
    "raster_blocks": "FP64 coverage test + L2 atomics",
    "raster_small": "FP64 coverage test + L2 atomics",
    "setup_raster": "FP64 coverage test + L2 atomics (small triangles are rasterised by the set-up thread)",
    "blur_h": "the recurrence (one multiply + seven dependent adds per sample and line)",
    "blur_v": "the recurrence (one multiply + seven dependent adds per sample and line)",
    "scan": "latency (single-CTA scan of a few thousand integers)",
}


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


def scene_of(workload, instance=None):
    """(scene path, assets dir, width, height) of a workload; generated instances are written on demand."""
    scene = WORKLOADS[workload][0]
    if scene.startswith("@"):
        from forkerrenderer_b200 import c5
        quads, w, h = c5.INSTANCES[instance or scene[1:]]
        root, path = c5.ensure(quads, w, h)
        return path, root, w, h
    path = os.path.join(REPO, scene)
    m = re.search(r"^screen\s+(\d+)\s+(\d+)", open(path).read(), re.M)
    return path, ASSETS, int(m.group(1)), int(m.group(2))


def config_of(workload, gpus):
    """The `config` object of the JSON line: identical for the product arm and the reference arm."""
    scene = WORKLOADS[workload][0]
    if scene.startswith("@"):
        from forkerrenderer_b200 import c5
        _, w, h = c5.INSTANCES[scene[1:]]
    else:
        m = re.search(r"^screen\s+(\d+)\s+(\d+)", open(os.path.join(REPO, scene)).read(), re.M)
        w, h = int(m.group(1)), int(m.group(2))
    return {"workload": WORKLOADS[workload][4], "width": w, "height": h, "gpus": gpus}


def reduced_scene(scene_path, factor):
    """Same scene at 1/factor of the linear resolution (bounded CPU sample).  Returns (path, width, height)."""
    txt = open(scene_path).read()
    m = re.search(r"^screen\s+(\d+)\s+(\d+)", txt, re.M)
    w, h = int(m.group(1)) // factor, int(m.group(2)) // factor
    txt = re.sub(r"^screen\s+\d+\s+\d+", "screen %d %d" % (w, h), txt, flags=re.M)
    fd, path = tempfile.mkstemp(suffix=".scene", prefix="fgl_bench_")
    os.write(fd, txt.encode())
    os.close(fd)
    return path, w, h


def time_reference(workload, frames, factor=1, instance=None):
    """Runs the unmodified reference on the scene (at 1/factor linear resolution, or on a smaller generated instance);
    returns per-frame seconds, the pixel size and a description of the sample."""
    _, shadow, wrap, filt, _ = WORKLOADS[workload]
    if not os.path.exists(REF_DRIVER):
        raise RuntimeError("oracle/_ref/ref_driver is missing: run `python -c 'import __graft_entry__ as g; g.build()'` where /root/reference exists")
    path, assets, w, h = scene_of(workload, instance)
    tmp = None
    if factor > 1:
        tmp, w, h = reduced_scene(path, factor)
    out = tempfile.mkdtemp(prefix="fgl_bench_ref_")
    r = subprocess.run([REF_DRIVER, "--assets", assets, "--scene", tmp or path, "--out", out, "--shadow", shadow, "--wrap", str(wrap),
                        "--filter", str(filt), "--frames", str(frames), "--quiet"], check=True, stdout=subprocess.PIPE, text=True)
    if tmp:
        os.unlink(tmp)
    times = [json.loads(l)["t_frame"] for l in r.stdout.splitlines() if l.startswith("{")]
    if instance:
        what = "the generated instance '%s' (%dx%d)" % (instance, w, h)
    elif factor > 1:
        what = "%s at %dx%d (1/%d linear resolution)" % (WORKLOADS[workload][0], w, h, factor)
    else:
        what = "%s at its full %dx%d" % (WORKLOADS[workload][0], w, h)
    return times, w, h, what


def reference_sample(workload, full):
    """How the CPU arm samples a workload: (factor, generated instance).  `full`: the --impl reference arm (minutes allowed)
    runs the scene files at their own resolution; the cpu_baseline leg of the default run stays within about 20 s."""
    if WORKLOADS[workload][0].startswith("@"):
        # a full C5 frame is ~10 minutes of reference time (load + 33 Mpixels of PCSS + SSAO): both legs use a smaller instance
        return 1, ("c5_golden" if full else "c5_cpu")
    if full:
        return 1, None
    return {"c3": 2, "c3_pbr": 2}.get(workload, 1), None


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    factor, instance = reference_sample(args.workload, True)
    # a full-resolution C3 frame is 35 - 70 s on one host core: at most two timed frames, no warm-up frame
    steps = max(1, min(args.steps, 2))
    times, w, h, what = time_reference(args.workload, steps, factor, instance)
    ms = 1e3 * sum(times) / len(times)
    mpx = w * h / 1e6 / (ms / 1e3)
    sample = "%s, %d frame(s) of oracle/_ref/ref_driver (the unmodified reference), single thread (the reference has no threads, forkergl.cpp:257)" % (what, len(times))
    line = {"impl": "reference", "metric": "mpixels_per_s", "value": mpx, "unit": "Mpixels/s", "n_gpus": args.gpus, "steps": len(times),
            "requested_steps": args.steps, "warmup": 0, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "reference assets (obj/*) / generated C5 OBJ, camera and light of the scene file",
            "config": config_of(args.workload, args.gpus),
            "frames_per_s": 1e3 / ms,
            "cpu_baseline": {"value": mpx, "unit": "Mpixels/s", "cores": 1, "kind": "reference", "sample": sample},
            "e2e": {"value": mpx, "unit": "Mpixels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def make_renderer(args, local, world, dist, instance=0):
    from forkerrenderer_b200 import binding as B
    from forkerrenderer_b200 import multigpu as M
    scene_file, shadow, wrap, filt, desc = WORKLOADS[args.workload]
    if scene_file.startswith("@") and instance == 0:
        # generated once per box: local rank 0 writes the files, the others wait for them
        if local == 0:
            scene_of(args.workload)
        if dist is not None:
            dist.barrier()
    path, assets, _, _ = scene_of(args.workload)
    host = B.product_host(instance)
    t0 = time.perf_counter()
    sc = host.load_scene(path, assets, wrap, filt)
    r = M.FacadeRenderer(host, sc, shadow, materialize=False)
    info = dict(desc=desc, triangles=sc.triangles, out_w=sc.width, out_h=sc.height, ssaa=bool(sc.ssaa), free=sc.free,
                load_s=time.perf_counter() - t0)
    return r, info


def _dbg(msg):
    if os.environ.get("FGL_BENCH_DEBUG"):
        print("[bench rank %s] %s" % (os.environ.get("RANK", "0"), msg), file=sys.stderr, flush=True)


def ncu_counters(workload, kernel):
    """Counters of the committed ncu capture for (workload, kernel), or None (profiles/r02_ncu_counters.json, r01_ncu_traffic.json)."""
    out = {}
    for name in ("r02_ncu_counters.json", "r01_ncu_traffic.json"):
        try:
            d = json.load(open(os.path.join(REPO, "profiles", name))).get(workload, {}).get(kernel)
        except Exception:
            d = None
        if isinstance(d, dict):
            for k, v in d.items():
                out.setdefault(k, v)
        elif d is not None:
            out.setdefault("dram_bytes_per_launch", d)
    return out or None


class Instance:
    """One frame in flight: its own facade instance (scene, fgl context, group membership), CUDA stream and host thread."""

    def __init__(self, args, local, world, rank, dist, index, torch, M):
        self.r, self.info = make_renderer(args, local, world, dist, instance=index)
        self.fgl, self.torch, self.world, self.rank = self.r.fgl, torch, world, rank
        self.stream = torch.cuda.Stream()
        self.fgl.set_stream(self.stream.cuda_stream)
        self.group = M.Group(self.fgl, dist, rank, world, self.r, mode=args.group) if world > 1 else None
        self.image = None
        # frames without host read-backs (hard / PCF shadows) go through frh_render_replay: recorded once as a CUDA graph, then
        # one launch per frame; PCSS frames (the default workloads) cannot be recorded and keep the eager calls
        self.replay = args.replay != "off" and world == 1 and WORKLOADS[args.workload][1] != "pcss"
        self.replayed = 0

    def frame(self, gather=True, eager=False):
        """One frame; with several GPUs: this rank's band (chain hand-off and gather of the 8-bit rows happen on the devices)."""
        with self.torch.cuda.stream(self.stream):
            if self.world > 1:
                return self.group.render_frame(gather=gather)
            if self.replay and not eager:
                self.replayed += int(self.r.replay())
                return None
            self.r.begin((0, -1))
            self.r.finish()
        return None

    def read(self):
        """The finished 8-bit frame in page-locked host memory (rank 0 of a group; other ranks wait for their stream)."""
        with self.torch.cuda.stream(self.stream):
            if self.world > 1:
                self.image = self.group.read_frame()
            else:
                self.image = self.fgl.read_plane("ssaa_u8" if self.info["ssaa"] else "frame_u8", pinned=True)
        return self.image

    def sync(self):
        self.fgl.sync()
        self.stream.synchronize()


def run_frames(insts, count, read, torch):
    """Renders `count` frames, frame i on instance i % len(insts), every instance driven by its own host thread (the frame
    call blocks in the two host read-backs of a PCSS frame, so one thread could not keep two frames in flight).
    Returns (device ms from the start to the last instance's end, CUDA events; wall-clock seconds)."""
    n = len(insts)
    for it in insts:
        it.sync()
    torch.cuda.synchronize()
    start = torch.cuda.Event(enable_timing=True)
    start.record(insts[0].stream)
    ends = [torch.cuda.Event(enable_timing=True) for _ in insts]
    errors = []
    device = torch.cuda.current_device()

    def work(k):
        try:
            torch.cuda.set_device(device)  # the current device is per thread
            it = insts[k]
            for _ in range(k, count, n):
                it.frame()
                if read:
                    it.read()
            ends[k].record(it.stream)
        except Exception as e:  # surfaced by the caller
            errors.append(e)

    t0 = time.perf_counter()
    if n == 1:
        work(0)
    else:
        threads = [threading.Thread(target=work, args=(k,)) for k in range(n)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
    if errors:
        raise errors[0]
    for it in insts:
        it.sync()
    wall = time.perf_counter() - t0
    return max(start.elapsed_time(e) for e in ends[: min(n, count)]), wall


def run_ours(args):
    import numpy as np
    import torch
    from forkerrenderer_b200 import binding as B  # noqa: F401
    from forkerrenderer_b200 import multigpu as M

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        # NCCL writes its debug output — at NCCL_DEBUG=VERSION / WARN that includes a version banner — to STDOUT unless told
        # otherwise; rank 0's stdout is the one JSON line (NCCL honours NCCL_DEBUG_FILE only above the VERSION level)
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    os.environ["FGL_DEVICE"] = str(local)

    # L2: a frame streams its planes (100 B / pixel) and, with SSAO / a stochastic shadow filter, its sample tables (384 B / pixel
    # of ball samples, >= 512 B / pixel of disk samples).  Where that is far beyond the 126 MB L2 nothing needs flushing; small
    # workloads get an explicit flush between frames and therefore render one frame at a time.
    scene_txt = "" if WORKLOADS[args.workload][0].startswith("@") else open(os.path.join(REPO, WORKLOADS[args.workload][0])).read()
    cfgd = config_of(args.workload, world)
    ssao_on = WORKLOADS[args.workload][0].startswith("@") or re.search(r"^ssao\s+on", scene_txt, re.M) is not None
    per_px = 100 + (384 if ssao_on else 0) + (512 if WORKLOADS[args.workload][1] in ("pcf", "pcss") else 0)
    k2 = 4 if re.search(r"^ssaa\s+on", scene_txt, re.M) else 1
    stream_bytes = cfgd["width"] * cfgd["height"] * k2 * per_px // world
    small = stream_bytes < (512 << 20)
    # default: frames in flight on one GPU (the latency-bound PCSS chain of one frame runs next to the other frames' passes, and a
    # frame's read-back next to the others' kernels): three while three contexts' sample tables and planes (about 1.5 kB per pixel
    # each) stay below a third of the HBM, else two; in a group the chain's serial relay through the bands bounds the frame, a
    # second frame in flight buys nothing (profiles/)
    ctx_bytes = cfgd["width"] * cfgd["height"] * k2 * 1500
    want = args.inflight if args.inflight > 0 else ((3 if 3 * ctx_bytes < (60 << 30) else 2) if world == 1 else 1)
    inflight = 1 if (small or (world > 1 and args.group != "peer")) else want

    _dbg("process group up, building %d renderer instance(s)" % inflight)
    insts = [Instance(args, local, world, rank, dist, k, torch, M) for k in range(inflight)]
    first = insts[0]
    fgl, info, group = first.fgl, first.info, first.group
    W, H = first.r.width, first.r.height            # raster size (output size x SSAA factor)
    out_px = info["out_w"] * info["out_h"]
    _dbg("group mode: %s" % (group.describe() if group else "single GPU"))

    def barrier():
        for it in insts:
            it.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda") if small else None

    def flush_l2(wait=True):
        """Overwrites 256 MiB (twice the L2) on the first instance's stream, in front of the next frame."""
        if flush_buf is not None:
            with torch.cuda.stream(first.stream):
                flush_buf.fill_(1)
            if wait:
                torch.cuda.synchronize()

    for _ in range(args.warmup):
        for it in insts:
            it.frame()
    barrier()
    _dbg("warm-up done")

    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.25)
    l0 = sum(it.fgl.launch_count() for it in insts)
    if flush_buf is None:
        ms_total, _ = run_frames(insts, args.steps, False, torch)
        ms = ms_total / args.steps
    else:  # one frame at a time, L2 flushed in between: flush, event, frame, event — all queued on one stream, so the events
        # bracket the frame's device work and not the host's way to the first launch (several GPUs: in step, frame by frame)
        ev = []
        for _ in range(args.steps):
            flush_l2(wait=world > 1)
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(first.stream)
            first.frame()
            e1.record(first.stream)
            ev.append((e0, e1))
        barrier()
        ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
    barrier()
    launches = sum(it.fgl.launch_count() for it in insts) - l0

    # ---- end to end: facade call(s) + the finished 8-bit frame in host memory, wall clock --------------------------
    frame_hash = None
    for it in insts:  # untimed: allocates the page-locked frame buffers, fingerprints the frame of every instance
        it.frame()
        img = it.read()
        if rank == 0:
            import hashlib
            h = hashlib.sha256(np.ascontiguousarray(img).tobytes()).hexdigest()
            assert frame_hash in (None, h), "the instances in flight rendered different frames"
            frame_hash = h
    barrier()
    tb0 = [it.fgl.transfer_bytes() for it in insts]
    if flush_buf is None:
        _, wall = run_frames(insts, args.steps, True, torch)
        e2e_ms = 1e3 * wall / args.steps
    else:
        e2e_t = []
        for _ in range(args.steps):
            flush_l2()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            first.frame()
            first.read()
            e2e_t.append(time.perf_counter() - t0)
        e2e_ms = 1e3 * sum(e2e_t) / len(e2e_t)
    tb1 = [it.fgl.transfer_bytes() for it in insts]
    h2d = sum(b[0] - a[0] for a, b in zip(tb0, tb1)) // args.steps
    d2h = sum(b[1] - a[1] for a, b in zip(tb0, tb1)) // args.steps
    if world > 1 and rank == 0:
        d2h += sum(it.group.host_read_bytes for it in insts) // max(1, len(insts))
    clocks = sampler.finish()
    _dbg("timed loops done: %.3f ms device, %.3f ms e2e" % (ms, e2e_ms))

    # ---- per-kernel breakdown (library instrumentation on ONE instance, separate SERIALISED frames: nothing overlaps while timing is on) ----
    barrier()
    fgl.enable_timing(True)
    fgl.reset_timings()
    nprof = 3
    for _ in range(nprof):
        flush_l2()
        if world > 1:
            dist.barrier()  # keep the ranks within one frame of each other
        first.frame(gather=False, eager=True)
        first.sync()
    kern = fgl.timings()
    fgl.enable_timing(False)
    fgl.reset_timings()
    for k in kern:
        k["ms_per_frame"] = k["ms_total"] / nprof
    kern.sort(key=lambda k: -k["ms_total"])
    peak, peak_src = measured_peaks()
    waits = ("pcss_peer_wait", "group_wait")   # spinning on another GPU is not this GPU's work
    work = [k for k in kern if k["name"] not in waits]
    top = work[0]
    t_launch = top["ms_total"] / top["launches"] / 1e3
    bytes_launch = top["algorithmic_bytes"] / top["launches"]
    achieved = bytes_launch / t_launch / 1e9
    total_kernel_ms = max(1e-9, sum(k["ms_per_frame"] for k in work))
    counters = ncu_counters(args.workload, top["name"])
    traffic = (counters or {}).get("dram_bytes_per_launch")
    roofline = {"kernel": top["name"], "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_launch,
                "avg_launch_ms": t_launch * 1e3, "share_of_frame": top["ms_per_frame"] / total_kernel_ms,
                "limiter": LIMITERS.get(top["name"], "HBM bandwidth"), "counters": counters,
                "note": "the kernel with the largest share of a frame's kernel time; `frac` = its algorithmic bytes over time against the HBM peak "
                        "(a small value next to a non-HBM `limiter` means the kernel is not a bandwidth problem); per-kernel table in `kernels`"}

    if world > 1:
        t = torch.tensor([ms, e2e_ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])

    _dbg("reporting")
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            factor, instance = reference_sample(args.workload, False)
            times, cw2, ch2, what = time_reference(args.workload, 1, factor, instance)
            cpu = {"value": cw2 * ch2 / 1e6 / times[0], "unit": "Mpixels/s", "cores": 1, "kind": "reference", "ms_per_frame": 1e3 * times[0],
                   "sample": "%s, 1 frame of oracle/_ref/ref_driver (the unmodified reference), single thread (the reference has no threads)" % what}
        line = {"metric": "mpixels_per_s", "value": out_px / 1e6 / (ms / 1e3), "unit": "Mpixels/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "reference assets (obj/*) / generated C5 OBJ, camera and light of the scene file",
                "config": config_of(args.workload, world),
                "details": {"triangles": info["triangles"], "scene_load_s": round(info["load_s"], 2),
                            "frames_in_flight": inflight,
                            "graph_replay": ("%d of the frames were launches of the recorded CUDA graph (frh_render_replay)" % sum(it.replayed for it in insts))
                                            if first.replay else "off (PCSS frames read chain state back to the host and cannot be recorded)",
                            "graph_replay_fallback": first.r.host.replay_fallback_reason() if first.replay else None,
                            "frames_in_flight_note": ("%d independent frames are rendered concurrently (one facade instance, fgl context, CUDA stream and host thread each); "
                                                      "ms_per_step is the time of K frames divided by K, the latency of a single frame is kernels[] summed" % inflight)
                                                     if inflight > 1 else "one frame at a time",
                            "partition": group.describe() if group else "single GPU",
                            "l2": "explicit 256 MiB flush between timed frames" if flush_buf is not None
                                  else "a frame streams %.0f MB per GPU (planes + sample tables), far beyond the 126 MB L2" % (stream_bytes / 1e6)},
                "frames_per_s": 1e3 / ms,
                "e2e": {"value": out_px / 1e6 / (e2e_ms / 1e3), "unit": "Mpixels/s", "ms_per_step": e2e_ms,
                        "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
                "gpu_launches": int(launches), "frame_sha256": frame_hash, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
                "single_frame_ms": round(sum(k["ms_per_frame"] for k in kern), 4),
                "kernels": [{"name": k["name"], "ms_per_frame": round(k["ms_per_frame"], 4), "launches_per_frame": k["launches"] / nprof,
                             "share": round(k["ms_per_frame"] / total_kernel_ms, 4) if k["name"] not in waits else None,
                             "gbps": (k["algorithmic_bytes"] / max(1e-12, k["ms_total"] / 1e3) / 1e9) if k["algorithmic_bytes"] else None,
                             "limiter": LIMITERS.get(k["name"])}
                            for k in kern]}
        print(json.dumps(line), flush=True)
    for it in insts:
        if it.group:
            it.group.close()
        it.info["free"]()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inflight", type=int, default=0,
                    help="frames rendered concurrently (each on its own facade instance / context / stream / host thread); 0 = default "
                         "(3 on one GPU up to 4K, 2 at 8K, 1 in a group); workloads small enough to need an L2 flush between frames always use 1")
    ap.add_argument("--replay", default="auto", choices=["auto", "off"],
                    help="auto: hard-shadow / PCF workloads on one GPU are rendered through frh_render_replay (the frame recorded once as a "
                         "CUDA graph, then one launch per frame); off: always the eager facade calls")
    ap.add_argument("--group", default="peer", choices=["peer", "nccl"],
                    help="N > 1: 'peer' = band-split passes exchanged by peer stores over NVLink, device-side flags (default); "
                         "'nccl' = replicated shadow pass, RGB8 bands all-gathered with NCCL, chain state through the host")
    args = ap.parse_args()
    if args.workload is None:
        # the metric is quoted per scene at 1 GPU (C3: the largest single-GPU configuration with a scene file) and, for the
        # scaling runs, on the north star's 8K / 10 M-triangle configuration
        args.workload = "c3" if args.gpus <= 1 else "c5"
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

"""Sort-first multi-GPU frames: one process per GPU, the frame split into horizontal row bands, geometry and textures
replicated, the finished 8-bit bands gathered over NCCL.

The only data dependency between bands is the PCSS sample-stream position (shadow.cpp:96-105): a band's chain starts
from the number of blockers found in all bands above it, so one integer travels rank r -> r + 1 between the two
halves of the frame (render_begin: shadow pass, raster, G-buffer, SSAO, blur; render_finish: chain + lighting)."""
import numpy as np


def band_rows(height, world, rank):
    """Rows [r0, r1) of rank `rank`; equal-sized bands (the last ones may be shorter or empty)."""
    per = (height + world - 1) // world
    r0 = min(height, rank * per)
    return r0, min(height, r0 + per), per


class FacadeRenderer:
    """A .scene file through the C++ facade (host.render_begin / render_finish)."""

    def __init__(self, host, scene, shadow_mode, materialize=False):
        self.host, self.scene, self.shadow_mode, self.materialize = host, scene, shadow_mode, materialize
        self.fgl = host.fgl
        self.width, self.height = scene.buffer_width, scene.buffer_height
        self.pcss = bool(scene.deferred and scene.shadow and shadow_mode in ("pcss", 2))

    def begin(self, band):
        self.fgl.set_row_band(*band)
        self.host.render_begin(self.scene, self.shadow_mode, self.materialize)

    def finish(self):
        self.host.render_finish(self.scene)

    def replay(self):
        """The whole frame through frh_render_replay (a recorded CUDA graph where the frame allows it).  True = graph launch."""
        self.fgl.set_row_band(0, -1)
        return self.host.render_replay(self.scene, self.shadow_mode, self.materialize)


class SyntheticRenderer:
    """A forkerrenderer_b200.synthetic.SyntheticScene through the raw C ABI."""

    def __init__(self, scene, width, height, **render_kw):
        self.scene, self.width, self.height, self.kw = scene, width, height, render_kw
        self.fgl = scene.f
        self.pcss = render_kw.get("shadow_mode") == 2 and not render_kw.get("forward") and render_kw.get("shadow", True)

    def begin(self, band):
        self.scene.render_begin(self.width, self.height, band=band, **self.kw)

    def finish(self):
        self.scene.render_finish()


def setup_peer_handoff(fgl, dist, rank, world, height):
    """Connects the contexts of a sort-first group for the DEVICE-side chain hand-off (fgl_chain_peer_*): every rank
    publishes the CUDA IPC handle of its mailbox, opens the next rank's, and from then on waits / signals inside its
    own stream.  Returns False (and leaves the host hand-off in place) if IPC is not available on some rank.  Empty bands
    (more GPUs than rows) take part like any other: their context waits for the count and passes it on."""
    if world < 2:
        return False
    ok = True
    try:
        _, handle = fgl.chain_peer_mailbox()
    except Exception:
        handle, ok = None, False
    handles = [None] * world
    dist.all_gather_object(handles, handle)
    if not ok or any(h is None for h in handles):
        return False
    try:
        fgl.chain_peer_connect(next_ipc=handles[rank + 1] if rank + 1 < world else None, wait_prev=rank > 0)
    except Exception:
        ok = False
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if not all(flags):
        fgl.chain_peer_connect(enable=False)
        return False
    return True


def render_frame(r, rank, world, comm=None, band_out=None, peer=False):
    """One frame of renderer `r` on this rank's band.  `comm` needs send_int(value, dst) / recv_int(src) when world > 1
    and the frame has a PCSS chain, unless the contexts were connected with setup_peer_handoff (peer=True: the chain
    state travels between the GPUs' streams without the host).  If band_out (a device pointer, bytes) is given, the
    band's RGB8 rows are copied there on the library's stream.  Returns (r0, r1)."""
    r0, r1, per = band_rows(r.height, world, rank)
    r.begin((r0, r1))
    if r.pcss and peer:
        r.finish()
    elif r.pcss:
        # every rank but the first receives, every rank but the last sends — also ranks whose band is empty (more GPUs than
        # rows): they pass the running count on unchanged, so that every send has its matching receive
        k = comm.recv_int(rank - 1) if world > 1 and rank > 0 else 0
        r.fgl.set_chain_blockers_before(k)
        r.finish()
        if world > 1 and rank < world - 1:
            comm.send_int(r.fgl.get_chain_blockers() if r1 > r0 else k, rank + 1)
    else:
        r.finish()
    if band_out is not None and r1 > r0:
        ptr, nbytes = band_out
        r.fgl.copy_plane_rows_to_device(11, r0, r1, ptr, (r1 - r0) * r.width * 3)  # FGL_PLANE_FRAME_RGB8
    return r0, r1


class TorchComm:
    """send / recv of one integer between ranks with torch.distributed (NCCL on the GPU box, gloo in the CPU tests).
    On CUDA the transfers run on their own stream: the chain total is known on the host while the band's filter and
    lighting kernels are still queued on the render stream, and must not wait behind them."""

    def __init__(self, dist, device):
        import torch
        self.dist, self.torch, self.device = dist, torch, device
        self.stream = torch.cuda.Stream(device=device) if getattr(device, "type", "cpu") == "cuda" else None

    def _ctx(self):
        import contextlib
        return self.torch.cuda.stream(self.stream) if self.stream is not None else contextlib.nullcontext()

    def send_int(self, value, dst):
        with self._ctx():
            t = self.torch.tensor([int(value)], dtype=self.torch.int64, device=self.device)
            self.dist.send(t, dst)

    def recv_int(self, src):
        with self._ctx():
            t = self.torch.zeros(1, dtype=self.torch.int64, device=self.device)
            self.dist.recv(t, src)
            return int(t.item())


def gather_bands(dist, torch, band_tensor, height, width, world):
    """all_gather of the equal-sized band buffers -> (height, width, 3) uint8 tensor on every rank."""
    per = band_tensor.shape[0]
    full = torch.empty((world * per, width, 3), dtype=torch.uint8, device=band_tensor.device)
    dist.all_gather_into_tensor(full, band_tensor)
    return full[:height]


class Group:
    """The sort-first group as bench.py drives it: one process per GPU; torch.distributed is only the rendezvous.

    mode 'peer' (default): the C++ group of include/forkergl_b200.h (fgl_group_*, csrc/group.cu).  The processes exchange
    their member records once (all_gather_object); from then on every frame is `frh_render` on each rank: band-split shadow
    pass and camera passes, shadow map / depth rows / finished 8-bit rows / PCSS chain state stored into the peers' memory
    by the producing kernels, ordered by device-side epoch flags.  Rank 0 reads the frame with frh_group_read_frame.

    mode 'nccl' (the round-1 path, kept for comparison): every GPU rasterises the whole shadow map and the whole depth
    plane, renders its band, the chain state travels through peer mailboxes (or NCCL send / recv) and the RGB8 bands are
    all-gathered with NCCL."""

    def __init__(self, fgl, dist, rank, world, renderer, mode="peer"):
        import torch
        self.torch, self.fgl, self.dist, self.rank, self.world, self.r, self.mode = torch, fgl, dist, rank, world, renderer, mode
        H, W = renderer.height, renderer.width
        self.r0, self.r1, self.per = band_rows(H, world, rank)
        self.host_read_bytes = 0
        if mode == "peer":
            host, scene = renderer.host, renderer.scene
            if not scene.deferred or scene.ssaa:
                raise RuntimeError("a sort-first group renders deferred frames without SSAA")
            mine = host.group_export(scene)
            members = [None] * world
            dist.all_gather_object(members, mine)
            host.group_connect(rank, world, members)
            dist.barrier()  # every context is connected before any of them begins a frame
            return
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.comm = TorchComm(dist, self.device)
        self.peer = bool(renderer.pcss and setup_peer_handoff(fgl, dist, rank, world, H))
        self.band = torch.empty((self.per, W, 3), dtype=torch.uint8, device=self.device)
        self.full = None
        self.host_frame = None

    def describe(self):
        if self.mode == "peer":
            return ("sort-first row bands, %d rows per GPU; shadow pass and camera passes band-split (per-triangle set-up culls to the band), shadow map rows / "
                    "depth rows / RGB8 rows / PCSS chain state stored into the peers' memory over NVLink by the producing kernels, device-side epoch flags, "
                    "no NCCL on the data path" % self.per)
        return ("sort-first row bands, %d rows per GPU, geometry and shadow pass replicated, RGB8 bands all-gathered with NCCL, PCSS chain state "
                "handed on through %s" % (self.per, "peer memory (device-side wait)" if self.peer else "the host (NCCL send/recv)"))

    def render_frame(self, gather=True):
        if self.mode == "peer":
            self.r.host.render(self.r.scene, self.r.shadow_mode, self.r.materialize)
            return None
        render_frame(self.r, self.rank, self.world, self.comm, band_out=(self.band.data_ptr(), self.band.numel()), peer=self.peer)
        if gather:
            self.full = gather_bands(self.dist, self.torch, self.band, self.r.height, self.r.width, self.world)
        return self.full

    def read_frame(self):
        """Rank 0: the finished frame in page-locked host memory (numpy); other ranks: waits for their stream."""
        if self.mode == "peer":
            if self.rank != 0:
                self.fgl.sync()
                return None
            img = self.r.host.group_read_frame(self.r.height, self.r.width)
            self.host_read_bytes = 0  # counted by the library (fgl_transfer_bytes)
            return img
        stream = self.torch.cuda.current_stream()
        if self.rank != 0:
            stream.synchronize()
            return None
        if self.host_frame is None:
            self.host_frame = self.torch.empty(self.full.shape, dtype=self.torch.uint8, pin_memory=True)
        self.host_frame.copy_(self.full, non_blocking=True)
        stream.synchronize()
        self.host_read_bytes = self.host_frame.numel()
        return self.host_frame.numpy()

    def close(self):
        if self.mode == "peer":
            self.fgl.sync()
            self.dist.barrier()
            self.r.host.group_disconnect()

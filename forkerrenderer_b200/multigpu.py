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
    own stream.  Returns False (and leaves the host hand-off in place) if a band would be empty or IPC is not available."""
    if world < 2 or band_rows(height, world, world - 1)[1] <= band_rows(height, world, world - 1)[0]:
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
        k = 0
        if world > 1 and rank > 0 and r1 > r0:
            k = comm.recv_int(rank - 1)
        r.fgl.set_chain_blockers_before(k)
        r.finish()
        if world > 1 and rank < world - 1:
            comm.send_int(r.fgl.get_chain_blockers() if r1 > r0 else 0, rank + 1)
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

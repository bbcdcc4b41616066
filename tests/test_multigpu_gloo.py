"""The N > 1 orchestration (band split, PCSS chain hand-off order, band gather) on CPU with the gloo backend,
world_size 2 and 3.  The per-band renderer is a stand-in built from one whole-frame oracle render: what is tested here
is the host-side logic of forkerrenderer_b200/multigpu.py, not the kernels (tests/test_gpu_synthetic.py checks that a
band render equals the corresponding rows of the full frame on the GPU)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import parity as P
from forkerrenderer_b200 import binding as B
from forkerrenderer_b200 import multigpu as M
from forkerrenderer_b200.synthetic import SyntheticScene


class FakeFgl:
    """Band-restricted view of a finished whole frame: rows of the RGB8 image + a per-row 'blocker' count."""

    def __init__(self, image, blockers_per_row):
        self.image, self.per_row = image, blockers_per_row
        self.before, self.band, self.log = None, None, []

    def set_row_band(self, r0, r1):
        self.band = (r0, r1)

    def set_chain_blockers_before(self, k):
        self.before = k
        self.log.append(("before", k))

    def get_chain_blockers(self):
        assert self.band[1] > self.band[0], "an empty band has no chain to ask"
        return self.before + int(self.per_row[self.band[0]:self.band[1]].sum())

    def copy_plane_rows_to_device(self, plane, r0, r1, ptr, nbytes):
        assert plane == B.PLANE_FRAME_RGB8 and nbytes == (r1 - r0) * self.image.shape[1] * 3
        ptr[: r1 - r0] = torch.from_numpy(self.image[r0:r1].copy())


class FakeRenderer:
    def __init__(self, fgl, pcss=True):
        self.fgl, self.pcss = fgl, pcss
        self.height, self.width = fgl.image.shape[:2]

    def begin(self, band):
        self.fgl.set_row_band(*band)

    def finish(self):
        assert not self.pcss or self.fgl.before is not None, "lighting started before the chain state arrived"


class PtrFgl(FakeFgl):
    pass


def worker(rank, world, port, image, per_row, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    fgl = FakeFgl(image, per_row)
    r = FakeRenderer(fgl)
    H, W = image.shape[:2]
    r0, r1, per = M.band_rows(H, world, rank)
    band = torch.zeros((per, W, 3), dtype=torch.uint8)
    comm = M.TorchComm(dist, "cpu")
    M.render_frame(r, rank, world, comm, band_out=(band, band.numel()))
    full = M.gather_bands(dist, torch, band, H, W, world)
    expected_before = int(per_row[:r0].sum())
    ok = fgl.before == expected_before and np.array_equal(full.numpy(), image)
    out.put((rank, ok, fgl.before, expected_before))
    dist.destroy_process_group()


class NoPeerFgl(FakeFgl):
    """A context whose library has no peer memory (the CPU oracle answers FGL_ERR_UNSUPPORTED the same way)."""

    def chain_peer_mailbox(self):
        raise B.FglError("no peer memory here")

    def chain_peer_connect(self, **kw):
        self.log.append(("peer_connect", kw))


class PeerFgl(FakeFgl):
    """Stand-in for a context with device-side hand-off: records what the driver asked for."""

    def chain_peer_mailbox(self):
        return 0x1000, bytes([7]) * 64

    def chain_peer_connect(self, **kw):
        self.log.append(("peer_connect", kw))


def peer_worker(rank, world, port, image, per_row, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    H = image.shape[0]
    # every rank must agree: one rank without IPC sends the whole group back to the host hand-off
    mixed = M.setup_peer_handoff(NoPeerFgl(image, per_row) if rank == 1 else PeerFgl(image, per_row), dist, rank, world, H)
    fgl = PeerFgl(image, per_row)
    ok = M.setup_peer_handoff(fgl, dist, rank, world, H)
    connect = [kw for name, kw in fgl.log if name == "peer_connect"]
    wired = (len(connect) == 1 and connect[0]["wait_prev"] == (rank > 0)
             and (connect[0]["next_ipc"] is None) == (rank == world - 1))
    # with the device-side hand-off the driver neither receives nor sends the chain state
    r = FakeRenderer(fgl, pcss=False)
    r.pcss = True
    r.finish = lambda: None

    class NoComm:
        def send_int(self, *a):
            raise AssertionError("host hand-off used although peer=True")
        recv_int = send_int
    M.render_frame(r, rank, world, NoComm(), peer=True)
    # a frame too short for the group (an empty last band) is wired like any other
    short = M.setup_peer_handoff(PeerFgl(image, per_row), dist, rank, world, world - 2)
    out.put((rank, (not mixed) and ok and wired and short))
    dist.destroy_process_group()


def test_peer_handoff_setup_and_fallback(oracle_fgl):
    image = np.zeros((12, 8, 3), dtype=np.uint8)
    per_row = np.ones(12, dtype=np.int64)
    world = 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=peer_worker, args=(r, world, port, image, per_row, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
    # the CPU oracle has no peer memory: the ABI says so instead of pretending
    with pytest.raises(B.FglError):
        oracle_fgl.chain_peer_mailbox()


class FakeGroupHost:
    """Stand-in for the facade's group entry points (frh_group_*): records what the Python driver asks for."""

    def __init__(self, rank):
        self.rank, self.log, self.fgl = rank, [], self

    def group_export(self, scene):
        self.log.append("export")
        return bytes([self.rank]) * 368

    def group_connect(self, rank, world, members, same_process=False):
        self.log.append(("connect", rank, world, [m[0] for m in members], len(members[0])))

    def render(self, scene, shadow_mode, materialize):
        self.log.append("render")

    def group_read_frame(self, h, w):
        self.log.append("read")
        return np.zeros((h, w, 3), dtype=np.uint8)

    def group_disconnect(self):
        self.log.append("disconnect")

    def sync(self):
        self.log.append("sync")


def group_worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Scene:
        deferred, ssaa = True, False

    class R:
        height, width, pcss, shadow_mode, materialize = 10, 8, True, "pcss", False
    host = FakeGroupHost(rank)
    R.host, R.scene = host, Scene()
    g = M.Group(host, dist, rank, world, R, mode="peer")
    g.render_frame()
    img = g.read_frame()
    g.close()
    connect = [e for e in host.log if isinstance(e, tuple)]
    ok = (host.log[0] == "export" and len(connect) == 1 and connect[0][1:4] == (rank, world, list(range(world))) and connect[0][4] == 368
          and host.log.count("render") == 1 and ((img is not None and "read" in host.log) if rank == 0 else (img is None and "read" not in host.log))
          and host.log[-1] == "disconnect" and "peers' memory" in g.describe())
    out.put((rank, ok, host.log if not ok else None))
    dist.destroy_process_group()


def test_peer_group_rendezvous_hands_every_rank_all_members_in_rank_order():
    """multigpu.Group(mode='peer'): the only host-side communication of the C++ / CUDA group is the exchange of the member
    records (rank order) and a barrier; rendering is frh_render on every rank, rank 0 alone reads the gathered frame."""
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=group_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _ in res), res


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,height", [(2, 50), (3, 50), (4, 6)])  # (4, 6): bands of 2 rows, the last band is empty
def test_band_split_handoff_and_gather(world, height, oracle_fgl):
    s = SyntheticScene(oracle_fgl)
    s.render(64, height, shadow_mode=B.SHADOW_PCSS)
    image = oracle_fgl.read_plane("frame_u8")
    rng = np.random.RandomState(1)
    per_row = rng.randint(0, 40, size=image.shape[0])
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, image, per_row, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    for rank, ok, before, expected in res:
        assert ok, (rank, before, expected)


def test_band_rows_cover_the_frame():
    for H in (1, 7, 800, 2160, 4320):
        for world in (1, 2, 3, 4, 8):
            rows = [M.band_rows(H, world, r) for r in range(world)]
            assert rows[0][0] == 0 and max(r[1] for r in rows) == H
            assert all(rows[i][1] == rows[i + 1][0] or rows[i + 1][0] == H for i in range(world - 1))
            assert all(r[1] - r[0] <= r[2] for r in rows)

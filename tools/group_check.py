#!/usr/bin/env python
"""Sort-first group parity on real GPUs, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 tools/group_check.py [workload ...]

Every rank first renders the whole frame alone, then the group renders it together (fgl_group_*: band-split passes, peer
stores, device-side flags).  Checked bit for bit: rank 0's gathered 8-bit frame against its own stand-alone frame, and on
every rank the band rows of the G-buffer / AO planes, the whole shadow map and the whole depth plane against its stand-alone
ones.  Three group frames are rendered (epoch flags, mailbox slots and plane reuse across frames)."""
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import bench  # noqa: E402


def main():
    import torch
    import torch.distributed as dist
    from forkerrenderer_b200 import binding as B
    from forkerrenderer_b200 import multigpu as M
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    os.environ["FGL_DEVICE"] = str(local)
    workloads = sys.argv[1:] or ["c1_ssao", "c1", "c5_small"]
    ok = True
    for wl in workloads:
        scene_file, shadow, wrap, filt, _ = bench.WORKLOADS[wl]
        if scene_file.startswith("@"):
            if local == 0:
                bench.scene_of(wl)
            dist.barrier()
        path, assets, _, _ = bench.scene_of(wl)
        host = B.product_host()
        sc = host.load_scene(path, assets, wrap, filt)
        fgl = host.fgl
        r = M.FacadeRenderer(host, sc, shadow, materialize=True)
        host.render(sc, shadow, True)
        names = ["frame_u8", "shadow", "depth", "normal", "worldpos", "lightndc", "albedo", "param", "ao", "frame"]
        alone = {n: fgl.read_plane(n).copy() for n in names}
        g = M.Group(fgl, dist, rank, world, r, mode="peer")
        bad = []
        for it in range(3):
            g.render_frame()
            img = g.read_frame()
            if rank == 0 and not np.array_equal(img, alone["frame_u8"]):
                d = np.abs(img.astype(int) - alone["frame_u8"].astype(int))
                bad.append("frame %d: gathered RGB8 differs (%d values, rows %s)" % (it, int((d > 0).sum()), np.unique(np.nonzero(d)[0])[:8]))
            fgl.sync()
            dist.barrier()
            for n in ("shadow", "depth"):
                got = fgl.read_plane(n)
                if not np.array_equal(got.view(np.uint32), alone[n].view(np.uint32)):
                    rows = np.unique(np.nonzero(got.view(np.uint32) != alone[n].view(np.uint32))[0])
                    bad.append("frame %d: plane %s differs on %d rows (%s...)" % (it, n, len(rows), rows[:6]))
            for n in ("normal", "worldpos", "lightndc", "albedo", "param", "ao", "frame"):
                got = fgl.read_plane(n)[g.r0:g.r1]
                if not np.array_equal(got.view(np.uint32), alone[n][g.r0:g.r1].view(np.uint32)):
                    bad.append("frame %d: band rows of %s differ" % (it, n))
            dist.barrier()
        g.close()
        sc.free()
        res = [None] * world
        dist.all_gather_object(res, bad)
        if rank == 0:
            flat = ["rank %d: %s" % (i, b) for i, bs in enumerate(res) for b in bs]
            print("%s on %d GPUs: %s" % (wl, world, "OK" if not flat else "FAILED\n  " + "\n  ".join(flat[:20])), flush=True)
            ok = ok and not flat
    ok_all = [None] * world
    dist.all_gather_object(ok_all, ok)
    dist.destroy_process_group()
    return 0 if all(o is not False for o in ok_all) and (rank != 0 or ok) else 1


if __name__ == "__main__":
    sys.exit(main())

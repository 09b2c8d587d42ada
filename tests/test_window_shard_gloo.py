"""Window-level sharding of one volume (parallel/window_shard.py): the plan, and -- with world_size-2/3 gloo processes on
CPU tensors -- the seam exchange and the slab gather.  The CUDA stitch kernel cannot run here, so each rank stitches its
slab with the oracle's stitcher restricted to the x-positions it holds (same z -> y -> x accumulation order); the
assembled map must be BIT-identical to the oracle's single-process stitch of all windows (utils.py:86-137)."""
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from light_unet.parallel.window_shard import exchange_seams, gather_slabs, plan_x_shards
from light_unet.utils import window_positions
from oracle import stitch_ref


@pytest.mark.parametrize("W,pw,ov,world", [(320, 48, 0.5, 8), (320, 48, 0.5, 2), (320, 48, 0.5, 3), (144, 48, 0.5, 4), (100, 48, 0.75, 3),
                                            (40, 48, 0.5, 2), (320, 96, 0.5, 8), (320, 64, 0.5, 16), (77, 16, 0.25, 5)])
def test_plan_covers_every_voxel_once(W, pw, ov, world):
    xpos = window_positions((W,), (pw,), ov)[0]
    shards = plan_x_shards(xpos, pw, W, world)
    assert sorted(c for s in shards for c in s.own) == list(range(len(xpos)))
    covered = np.zeros(W, dtype=int)
    for s in shards:
        if not s.own:
            assert s.x0 == s.x1 == 0 and not s.need and not s.recv
            continue
        covered[s.x0:s.x1] += 1
        # every window that covers a voxel of the slab is held: computed here or received from an EARLIER rank
        for x in range(s.x0, s.x1):
            cover = [c for c in range(len(xpos)) if xpos[c] <= x < xpos[c] + pw]
            assert set(cover) <= set(s.need), (s.rank, x)
        assert all(src < s.rank for src in s.recv)
        assert s.need == sorted(s.need) and s.need[len(s.need) - len(s.own):] == s.own
    assert (covered == 1).all()
    for s in shards:                      # send lists mirror receive lists
        for dst, cols in s.send.items():
            assert shards[dst].recv[s.rank] == cols


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _slab_stitch(shape, patch, zpos, ypos, xl, local, x0, x1, gauss):
    """The oracle's stitcher over the x-positions `xl` held by one rank (`local`: [len(xl)][nz][ny] windows), cropped to the
    voxel slab [x0, x1)."""
    nz, ny = len(zpos), len(ypos)
    preds = np.stack([local[(c * nz + a) * ny + b] for a in range(nz) for b in range(ny) for c in range(len(xl))])
    return stitch_ref.stitch(shape, patch, (zpos, ypos, xl), preds, gauss)[:, :, x0:x1]


def _worker(rank, world, port, shape, patch, ov, ret):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        zpos, ypos, xpos = window_positions(shape, patch, ov)
        nz, ny, nx = len(zpos), len(ypos), len(xpos)
        rng = np.random.default_rng(7)                                     # same "predictions" on every rank
        allp = rng.random((nz * ny * nx,) + tuple(patch), dtype=np.float32)     # z -> y -> x window order
        shards = plan_x_shards(xpos, patch[2], shape[2], world)
        me = shards[rank]
        per = int(np.prod(patch))
        local = torch.full((len(me.need) * nz * ny, per), float("nan"))
        for j, c in enumerate(me.own):                                     # "compute" the own positions
            for a in range(nz):
                for b in range(ny):
                    local[((me.n_recv + j) * nz + a) * ny + b] = torch.from_numpy(allp[(a * ny + b) * nx + c].ravel())
        if me.need:
            exchange_seams(local.view(len(me.need), nz * ny * per), me, None)
        assert not torch.isnan(local).any()
        full = torch.zeros(shape) if rank == 0 else None
        slab = None
        if me.own:
            got = _slab_stitch(shape, patch, zpos, ypos, [xpos[c] for c in me.need], local.numpy().reshape((-1,) + tuple(patch)),
                               me.x0, me.x1, True)
            slab = torch.from_numpy(np.ascontiguousarray(got))
            if rank == 0:
                full[:, :, me.x0:me.x1] = slab
        gather_slabs(full, slab, shards, rank, None)
        if rank == 0:
            want = stitch_ref.stitch(shape, patch, (zpos, ypos, xpos), allp, True)
            ret["equal"] = bool(np.array_equal(full.numpy(), want))
            ret["maxdiff"] = float(np.abs(full.numpy() - want).max())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,shape,patch", [(2, (20, 20, 100), (16, 16, 16)), (3, (16, 24, 77), (16, 16, 16)), (2, (12, 12, 40), (16, 16, 48))])
def test_sharded_stitch_is_bit_identical(world, shape, patch):
    port = _free_port()
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, shape, patch, 0.5, ret), nprocs=world, join=True)
        assert ret["equal"], ret["maxdiff"]

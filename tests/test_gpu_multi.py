"""Multi-GPU parity (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu -s`; skipped on a
one-GPU box).  One process per GPU over NCCL, exactly as bench.py runs:

  * DataParallelStep with the real model and the CUDA Focal Tversky loss: a dp-2 step on two half batches == the
    single-process step on the concatenated batch (loss, every parameter gradient, updated parameters) -- the Tversky sums
    are all-reduced before the ratio and the gradients reduced with SUM (trainer.py:223-232, losses.py:44-49);
  * window-level sharding of one volume (parallel/window_shard.py, utils.py:86-137): bit-identical to the single-GPU
    stitch given identical window predictions, and equal to round-off end to end (map, mask, boxes).
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import synth, unet_ref

pytestmark = pytest.mark.gpu


def _need_gpus(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model(dtype, dropout_p, dev, wseed=1):
    from light_unet.models import Lightweight3DUNet
    cfg = unet_ref.UNetCfg(dropout_p=dropout_p)
    sd_np = synth.synth_state_dict(unet_ref.param_shapes(cfg), wseed)
    m = Lightweight3DUNet(dropout_p=dropout_p)
    m.load_state_dict(unet_ref.to_torch(sd_np))
    return m.to(dev).set_compute_dtype(dtype)


def _dp_worker(rank, world, port, dtype, ret):
    from light_unet.models import FocalTverskyLoss
    from light_unet.parallel import DataParallelStep
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    try:
        B, S = 4, 24
        x, t = synth.synth_patches(world * B, (S, S, S), 42)
        xs, ts = torch.from_numpy(x).to(dev), torch.from_numpy(t).to(dev)
        model = _model(dtype, 0.0, dev).train()
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5)
        stepper = DataParallelStep(model, FocalTverskyLoss(), opt, world_size=world)
        loss = stepper.step(xs[rank * B:(rank + 1) * B], ts[rank * B:(rank + 1) * B])
        grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
        torch.cuda.synchronize()
        if rank == 0:
            ref = _model(dtype, 0.0, dev).train()
            ropt = torch.optim.AdamW(ref.parameters(), lr=1e-4, weight_decay=1e-5)
            rloss = DataParallelStep(ref, FocalTverskyLoss(), ropt, world_size=1).step(xs, ts)
            rgrads = {k: p.grad.detach() for k, p in ref.named_parameters()}
            # Per-tensor rel-L2 against the single-process step (floor: 0.1 % of the largest gradient norm).  The kernels
            # combine every run-order-dependent partial sum in double, so a sample's forward tensors -- and with them every
            # LeakyReLU / max-pool decision of its backward pass -- are the same whichever batch it sits in; what is left
            # between 4 + 4 samples all-reduced and 8 in one launch is the fp32 summation order of the weight gradients
            # (measured 2e-6).  A wrong reduction (mean instead of sum, Tversky ratio taken per rank) would show up as O(1)
            # differences in every tensor and in the loss.  [Accuracy against the float64 oracle is the single-GPU tests'
            # subject (test_gpu_configs.py): per-tensor it is dominated by rare LeakyReLU sign flips between precisions --
            # ~1e-3 of a downstream gradient tensor each -- which hit this batch's third sample in fp32-storage mode.]
            gmax = max(float(v.norm()) for v in rgrads.values())
            errs = {}
            for k in grads:
                d, n = float((grads[k] - rgrads[k]).norm()), float(rgrads[k].norm())
                errs[k] = d / max(n, 1e-3 * gmax)
            worst = max(errs, key=errs.get)
            perr = max(float((p.detach() - q.detach()).abs().max()) for p, q in zip(model.parameters(), ref.parameters()))
            ret.update(loss=float(loss), rloss=float(rloss), worst=worst, werr=errs[worst], perr=perr)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype,tol", [("f32", 1e-4), ("f16", 1e-4)])
def test_data_parallel_step_nccl_equals_single_process(dtype, tol):
    _need_gpus(2)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_dp_worker, args=(2, _free_port(), dtype, ret), nprocs=2, join=True)
        print(f"dp2/{dtype}: loss {ret['loss']:.7f} vs single-process {ret['rloss']:.7f}; worst per-tensor gradient rel-L2 "
              f"{ret['werr']:.3e} ({ret['worst']}); max |param - param_single| after AdamW {ret['perr']:.3e}")
        assert abs(ret["loss"] - ret["rloss"]) < 1e-5
        assert ret["werr"] < tol and ret["perr"] < 1e-5            # same gradient signs: AdamW's first step moves both by the same +-lr


def _shard_worker(rank, world, port, ret):
    from light_unet.core.inferencer import Inferencer
    import light_unet.utils as lu
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world, device_id=dev)
    try:
        inf = Inferencer.__new__(Inferencer)
        inf.config = {"data": {"bbox_expansion_voxels": 3, "patch_size": [48, 48, 48], "volume_threshold": {"inference_cc": 0.5}}}
        inf.device = dev
        inf.model = _model("f32", 0.0, dev, wseed=3).eval()
        vol_np = synth.synth_volume((64, 56, 200), seed=5, n_blobs=3)
        vol = torch.from_numpy(vol_np).to(dev)
        out = {}
        # (a) end to end: sharded == single GPU up to the round-off of two forward passes
        prob_s, boxes_s = inf.infer_volume(vol, threshold=0.5, return_device=True, shard=(rank, world, None))
        if rank == 0:
            prob_1, boxes_1 = inf.infer_volume(vol, threshold=0.5, return_device=True)
            out["e2e_maxdiff"] = float((prob_s - prob_1).abs().max())
            # the two maps differ by the round-off of two forward passes (different window batches -> different summation
            # order of the InstanceNorm statistics), so a voxel within that distance of the threshold may flip; the box lists
            # must be equal whenever no voxel flipped, and every flip must lie inside the round-off band
            flips = (prob_s > 0.5) != (prob_1 > 0.5)
            out["nflips"] = int(flips.sum())
            out["flips_in_band"] = bool(((prob_1[flips] - 0.5).abs() <= 1e-5).all()) if out["nflips"] else True
            out["boxes_equal"] = [b["bbox_voxel"] for b in boxes_s] == [b["bbox_voxel"] for b in boxes_1]
            out["nboxes"] = len(boxes_1)
        # (b) identical predictions (a deterministic function of the window position) -> bit-identical maps
        orig = lu._forward_windows

        def fake(model, v, pos_d, nwin, patch, preds, window_batch):
            p = pos_d[:nwin].to(torch.float32)
            base = ((p[:, 0] * 0.013 + p[:, 1] * 0.0071 + p[:, 2] * 0.0037) % 1.0).view(-1, 1, 1, 1, 1)
            ramp = torch.linspace(0, 1, patch[0] * patch[1] * patch[2], device=v.device).view(1, 1, *patch)
            preds[:nwin] = (base * 0.5 + ramp * 0.5).to(torch.float32)
        lu._forward_windows = fake
        try:
            ps, _ = lu.sliding_window_device(vol, inf.model, (48, 48, 48), 0.5, True, shard=(rank, world, None))
            if rank == 0:
                p1, _ = lu.sliding_window_device(vol, inf.model, (48, 48, 48), 0.5, True)
                out["bit_identical"] = bool(torch.equal(ps, p1))
        finally:
            lu._forward_windows = orig
        torch.cuda.synchronize()
        if rank == 0:
            ret.update(out)
    finally:
        dist.destroy_process_group()


def test_window_sharded_volume_nccl():
    _need_gpus(2)
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_shard_worker, args=(2, _free_port(), ret), nprocs=2, join=True)
        print(f"window-sharded x2: bit-identical given identical predictions: {ret['bit_identical']}; end to end max |diff| "
              f"{ret['e2e_maxdiff']:.3e}, {ret['nboxes']} boxes, box lists equal: {ret['boxes_equal']}, threshold flips: {ret['nflips']} "
              f"(all within 1e-5 of the threshold: {ret['flips_in_band']})")
        assert ret["bit_identical"] and ret["e2e_maxdiff"] < 1e-5 and ret["flips_in_band"]
        assert ret["boxes_equal"] or ret["nflips"] > 0

#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200-native Light-3D-Unet hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--variant dws|grouped|dense] [--dtype f16|f32] [--train-dtype f32|f16] [--skip-train] [--skip-cpu]
                    [--sweep] [--patch P --batch B] [--eager-gpu]

One JSON line on stdout (rank 0).  Metric (BASELINE.json): sliding-window inference volume-voxels/s on the
synthetic whole-body 4 mm PET volume (128x128x320, 48^3 windows, 50 % overlap, Gaussian stitching, threshold
-> connected components -> bounding boxes); the 48^3-patch training step (fwd + Focal Tversky + bwd + AdamW,
batch 8 per GPU) is reported in the same line under "train".

A "step" of the headline metric is one whole volume through the hot path.
  value : volume-voxels/s with the volume already resident in HBM (device-timed, CUDA events, max over ranks)
  e2e   : the same through the reference-facing API, Inferencer.infer_volume(host array): pinned H2D copy of the
          volume and D2H read of the probability map + box table inside the timed region
N > 1: one process per GPU (torchrun), volumes are independent -> every rank runs its own volumes, no data-path
collective ("weak" scaling); training is data parallel with the gradient all-reduce over NCCL.  The same line also
carries "latency": ONE volume whose windows are sharded over the N ranks (seam exchange + slab gather over NCCL,
parallel/window_shard.py) -- single-volume latency, strong scaling.
--sweep (configs[4]): forward throughput for patch 48^3 / 64^3 / 96^3 x batch 1..64 per GPU, under "sweep".
--patch P --batch B: one point of that sweep as the headline workload instead of the volume.
Storage modes: inference is quoted in fp16 storage (1e-2 bar met with margin, tests/test_gpu_configs.py); the training
step in fp32 storage, the mode whose every parameter gradient matches the reference (--train-dtype f16 for the faster,
direction-only mode).
--impl reference: the oracle's CPU restatement of the reference path (torch-CPU ATen kernels, all host threads)
on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "light-3d-unet-front_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np   # noqa: E402
import torch         # noqa: E402

VOLUME = (128, 128, 320)
PATCH = (48, 48, 48)
NVOX = VOLUME[0] * VOLUME[1] * VOLUME[2]
NWIN = 5 * 5 * 13
TRAIN_BATCH = 8
VARIANTS = {"dws": dict(use_depthwise_separable=True, use_grouped=True),
            "grouped": dict(use_depthwise_separable=False, use_grouped=True),
            "dense": dict(use_depthwise_separable=False, use_grouped=False)}
# compulsory activation elements per 48^3 patch (each tensor crossing an InstanceNorm
# reduction written once + read once) and forward FLOP per patch
ELEMS_PER_PATCH = 21.32e6
FWD_FLOP = {"dws": 1.402e9, "grouped": 3.354e9, "dense": 12.47e9}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback"}


def measured_traffic(kernel: str):
    """DRAM bytes per launch of the dominant entry point (dram__bytes_read.sum + dram__bytes_write.sum) from the committed
    ncu capture of this workload (profiles/traffic.json, measured offline: numbers taken under a profiler are never timed)."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return int(json.load(f)[kernel]["bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region: NVML in-process every 2 ms (a timed region of a few
    steps is tens of milliseconds -- shorter than one `nvidia-smi -lms` period); nvidia-smi every 200 ms as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None
        self.thread, self.stop_flag, self.samples, self.bits, self.max_mhz = None, False, [], 0, None
        self.nvml, self.handle = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.nvml, self.handle = pynvml, h
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nvml = None

    def _poll(self):
        nv, h = self.nvml, self.handle
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                try:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(h))
                except Exception:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nvml is not None:
            import threading
            self.stop_flag, self.samples, self.bits = False, [], 0
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            self.thread = None
            if self.samples:
                out.update(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.max_mhz,
                           reasons=sorted(n for b, n in self.REASONS if self.bits & b), samples=len(self.samples), source="nvml")
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, mx = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 7:
                    continue
                sm.append(float(f[0]))
                mx = float(f[1])
                for nme, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm), source="nvidia-smi")
        return out


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def max_over_ranks(ms: float, world: int, device) -> float:
    if world == 1:
        return ms
    import torch.distributed as dist
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier(world):
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def timed(fn, steps, world, device):
    """barrier+sync | K x fn on the current stream between two CUDA events | sync+barrier -> max-over-ranks ms"""
    barrier(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    barrier(world)
    return max_over_ranks(ms, world, device)


# ------------------------------------------------------------------------------------------ ours
def make_inferencer(variant, dtype, device):
    from light_unet.core.inferencer import Inferencer
    from oracle import synth, unet_ref   # synthetic weights only (shared seeded generator); not on the timed path
    kw = VARIANTS[variant]
    cfg = unet_ref.UNetCfg(dropout_p=0.0, **kw)
    sd = unet_ref.to_torch(synth.synth_state_dict(unet_ref.param_shapes(cfg), 3))
    d = tempfile.mkdtemp(prefix="l3d_bench_")
    ckpt = os.path.join(d, "model.pth")
    torch.save({"epoch": 0, "model_state_dict": sd, "best_epoch": 0, "best_metric": 0.0}, ckpt)
    config = {"model": {"output_channels": 1, "start_channels": 16, "encoder_channels": [16, 32, 64, 128],
                        "use_depthwise_separable": kw["use_depthwise_separable"], "use_grouped_conv": kw["use_grouped"],
                        "groups": 8},
              "output": {"prob_maps_dir": os.path.join(d, "prob"), "bboxes_dir": os.path.join(d, "bbox")},
              "data": {"patch_size": list(PATCH), "bbox_expansion_voxels": 3, "volume_threshold": {"inference_cc": 0.5}},
              "validation": {"default_threshold": 0.3}}
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        inf = Inferencer(config, ckpt)
    inf.model.set_compute_dtype(dtype)
    return inf


def kernel_table(nv, rec):
    """TIMER records -> (per-kernel, per-entry-point) {name: [launches, ms, algorithmic bytes]}: dispatching entry points
    (l3d_dwpw_fwd ...) are attributed to the kernel they launched, the others launch one kernel of their own name."""
    by_name = {}
    for (name, tag), (n, t, b) in rec.items():
        a = by_name.setdefault(name, [0, 0.0, 0])
        a[0] += n; a[1] += t; a[2] += b
    by_kernel = dict(nv.TIMER.by_kernel)
    for k, v in by_name.items():
        if k not in nv._DISPATCHING:
            by_kernel[k] = tuple(v)
    return by_kernel, by_name


def bench_patches(args, inf, device, world, patch, batch, steps, warmup):
    """configs[4]: forward of `batch` patches of patch^3 per GPU through the nn.Module (inputs resident, rotating over
    enough distinct batches to exceed L2)."""
    from oracle import synth
    nb = max(2, min(8, int(np.ceil(300e6 / (batch * patch ** 3 * 4)))))
    xs = [torch.from_numpy(synth.synth_patches(batch, patch, seed=7 + i)[0]).to(device) for i in range(nb)]
    model = inf.model.eval()

    def step(i):
        with torch.no_grad():
            model(xs[i % nb])
    for i in range(warmup):
        step(i)
    ms = timed(step, steps, world, device)
    es = 2 if args.dtype == "f16" else 4
    scale = (patch / 48.0) ** 3
    per_s = world * batch * steps / (ms * 1e-3)
    return {"patch": patch, "batch_per_gpu": batch, "ms_per_step": round(ms / steps, 4), "patches_per_s": round(per_s, 1),
            "patch_voxels_per_s": round(per_s * patch ** 3, 1),
            "frac_hbm": round(per_s / world * scale * ELEMS_PER_PATCH * es * 2 / 1e9 / peaks()["hbm_gbs"], 4)}


def bench_ours(args):
    from light_unet import _native as nv
    from oracle import synth
    rank, world, local = dist_env()
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
    inf = make_inferencer(args.variant, args.dtype, device)
    if args.patch:
        # one point of the configs[4] sweep as the headline workload
        l0 = nv.launch_count()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        r = bench_patches(args, inf, device, world, args.patch, args.batch, args.steps, args.warmup)
        clocks = sampler.stop() if rank == 0 else None
        if rank == 0:
            emit({"metric": "patch forward patch-voxels/s", "value": r["patch_voxels_per_s"], "unit": "voxels/s", "n_gpus": world,
                  "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                  "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
                  "config": {"workload": f"configs[4]: forward of {args.batch} patches of {args.patch}^3 per GPU through Lightweight3DUNet.forward",
                             "variant": args.variant, "patch": args.patch, "batch_per_gpu": args.batch,
                             "l2": "inputs rotate over enough distinct batches to exceed the 126 MB L2"},
                  "patches_per_s": r["patches_per_s"], "frac_hbm": r["frac_hbm"], "gpu_launches": int(nv.launch_count() - l0), "clocks": clocks})
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    # 8 distinct volumes (168 MB > the 126 MB L2) rotated between steps; every step also streams > 10 GB of
    # activations through HBM, so no input survives in L2 from one step to the next
    base = synth.synth_volume(VOLUME, seed=42 + rank, n_blobs=6)
    host_vols = [torch.from_numpy(np.roll(base, 7 * i, axis=2).copy()).pin_memory() for i in range(8)]
    dev_vols = [h.to(device) for h in host_vols]
    nboxes = [0]

    def step_resident(i):
        prob, boxes = inf.infer_volume(dev_vols[i % 8], threshold=0.3, return_device=True)
        nboxes[0] = len(boxes)

    host_out = torch.empty(VOLUME, dtype=torch.float32).pin_memory()

    def step_e2e(i):
        # host volume in (pinned), host probability map + box list out: the D2H copy of the map runs on a side stream
        # under the connected-component / box kernels and is complete when infer_volume returns
        prob, boxes = inf.infer_volume(host_vols[i % 8], threshold=0.3, prob_out=host_out)
        nboxes[0] = len(boxes)

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = nv.launch_count()
    ms = timed(step_resident, args.steps, world, device)
    launches = nv.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    for i in range(max(1, args.warmup // 2)):
        step_e2e(i)
    ms_e2e = timed(step_e2e, args.steps, world, device)
    # the same through the streaming form of the public API: Inferencer.infer_volumes uploads volume i + 1 (pinned host ->
    # device) and downloads map i - 1 + its box table on side streams under the kernels of volume i; every step still copies
    # its 21 MB in and its 21 MB map out, and every result is read on the host (box list built) inside the timed region
    outs2 = [torch.empty(VOLUME, dtype=torch.float32).pin_memory() for _ in range(3)]
    stream_steps = args.steps + 3          # two primed + `steps` timed, each of which launches one volume and completes one; one spare
    gen = inf.infer_volumes((host_vols[i % 8] for i in range(stream_steps)), threshold=0.3, prob_outs=(outs2[i % 3] for i in range(stream_steps)))
    next(gen); next(gen)                                                 # pipeline primed: 3 volumes launched, 2 completed (untimed)

    def step_stream(i):
        prob, boxes = next(gen)
        nboxes[0] = len(boxes)
    ms_stream = timed(step_stream, args.steps, world, device)           # every timed call: stage(i+2), launch(i+1), complete(i)
    gen.close()

    # per-kernel device time (CUDA events around every libl3d launch, same stream) over one more volume
    nv.TIMER.start()
    step_resident(0)
    rec = nv.TIMER.stop()
    tot_ms = sum(v[1] for v in rec.values())
    by_kernel, by_name = kernel_table(nv, rec)
    top = max(by_kernel.items(), key=lambda kv: kv[1][1])
    pk = peaks()
    es = 2 if args.dtype == "f16" else 4
    achieved = top[1][2] / (top[1][1] * 1e-3) / 1e9 if top[1][1] > 0 else 0.0
    step_ms = ms / args.steps
    roofline = {"bound": "hbm", "kernel": top[0], "achieved": round(achieved, 1), "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": round(achieved / pk["hbm_gbs"], 4), "traffic": measured_traffic(top[0]), "peak_source": pk["source"],
                "launches_per_step": top[1][0], "avg_launch_us": round(1e3 * top[1][1] / top[1][0], 1),
                "share_of_step": round(top[1][1] / tot_ms, 3),
                "algorithmic_bytes_per_launch": int(top[1][2] / top[1][0]),
                # whole-network view: compulsory activation bytes per volume (21.32 M materialised elements per
                # 48^3 patch, each written once and read once = 85.3 MB in 16-bit storage; 325 windows = 27.7 GB) / step time;
                # "declared" = the sum of the per-launch algorithmic bytes the engine declares for the kernels it really runs
                "network": {"algorithmic_gb_per_step": round(NWIN * ELEMS_PER_PATCH * es * 2 / 1e9, 2),
                            "achieved_gbs": round(NWIN * ELEMS_PER_PATCH * es * 2 / (step_ms * 1e-3) / 1e9, 1),
                            "frac_hbm": round(NWIN * ELEMS_PER_PATCH * es * 2 / (step_ms * 1e-3) / 1e9 / pk["hbm_gbs"], 4),
                            "declared_gb_per_step": round(sum(v[2] for v in by_kernel.values()) / 1e9, 2),
                            "declared_frac_hbm": round(sum(v[2] for v in by_kernel.values()) / (step_ms * 1e-3) / 1e9 / pk["hbm_gbs"], 4),
                            "tflops": round(NWIN * FWD_FLOP[args.variant] / (step_ms * 1e-3) / 1e12, 2),
                            "frac_tensor": round(NWIN * FWD_FLOP[args.variant] / (step_ms * 1e-3) / 1e12 / pk["bf16_tflops"], 4)},
                "kernels": {k: {"launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / tot_ms, 3),
                                "gbs": round(v[2] / max(v[1], 1e-9) / 1e6, 1)} for k, v in
                            sorted(by_kernel.items(), key=lambda kv: -kv[1][1])},
                "entry_points": {k: {"launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / tot_ms, 3)} for k, v in
                                 sorted(by_name.items(), key=lambda kv: -kv[1][1])}}

    # ---- single-volume latency with the windows of ONE volume sharded over the ranks (strong scaling)
    latency = None
    if world > 1 and not args.no_latency:
        shard = (rank, world, None)
        shared = torch.from_numpy(synth.synth_volume(VOLUME, seed=42, n_blobs=6)).to(device)     # the same volume on every rank

        def step_sharded(i):
            inf.infer_volume(shared, threshold=0.3, return_device=True, shard=shard)
        for i in range(3):
            step_sharded(i)
        ms_lat = timed(step_sharded, args.steps, world, device)
        latency = {"workload": "ONE 128x128x320 volume, its 13 x-positions of windows sharded over the ranks "
                               "(seam exchange + slab gather over NCCL, threshold -> CC -> boxes on rank 0)",
                   "ms_per_volume": round(ms_lat / args.steps, 3), "value": round(NVOX * args.steps / (ms_lat * 1e-3), 1),
                   "unit": "voxels/s", "scaling": "strong", "n_gpus": world, "ms_per_volume_1gpu_volume_sharded": round(step_ms, 3)}

    sweep = None
    if args.sweep:
        sweep = []
        for patch in (48, 64, 96):
            for batch in (1, 2, 4, 8, 16, 32, 64):
                if batch * patch ** 3 > 64 * 64 ** 3 * 2:          # 96^3 x 64 (57 M voxels, ~36 GB fp16 workspace) is skipped
                    continue
                sweep.append(bench_patches(args, inf, device, world, patch, batch, max(3, args.steps), 3))
                inf.model._plan.drop_workspaces()

    train = None
    if not args.skip_train:
        inf.model._plan.drop_workspaces()
        train = bench_train(args, device, rank, world, args.train_dtype)
        if args.train_dtype != "f16" and not args.quick:
            train["fp16_storage"] = bench_train(args, device, rank, world, "f16")

    eager = None
    if args.eager_gpu and rank == 0:
        eager = eager_gpu_leg(args.variant, device)

    cpu = None
    if rank == 0 and not args.skip_cpu:
        cpu = cpu_baseline(args.variant, target_s=12.0)

    if rank == 0:
        line = {"metric": "sliding-window inference volume-voxels/s", "value": round(world * NVOX * args.steps / (ms * 1e-3), 1),
                "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": round(step_ms, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.dtype, "data": "synthetic",
                "config": {"workload": "configs[2]: sliding-window inference, synthetic whole-body PET 128x128x320 @4mm, "
                                       "48^3 windows, 50% overlap (325 windows), Gaussian stitch, threshold 0.3 -> CC -> bbox",
                           "variant": args.variant, "params": {"dws": 217228, "grouped": 391521, "dense": 2308737}[args.variant],
                           "volumes_per_step_per_gpu": 1, "window_batch": window_batch(),
                           "l2": "inputs rotate over 8 volumes (168 MB) and each step streams >10 GB of activations (> 126 MB L2)",
                           "parallelism": f"volume-sharded x{world}", "boxes_found": nboxes[0]},
                "e2e": {"value": round(world * NVOX * args.steps / (ms_stream * 1e-3), 1), "unit": "voxels/s",
                        "ms_per_step": round(ms_stream / args.steps, 3),
                        "api": "Inferencer.infer_volumes(host volumes): upload of volume i+1 and download of map i-1 overlapped with volume i",
                        "one_call_per_volume": {"value": round(world * NVOX * args.steps / (ms_e2e * 1e-3), 1),
                                                "ms_per_step": round(ms_e2e / args.steps, 3), "api": "Inferencer.infer_volume(host volume)"},
                        "h2d_bytes_per_step": NVOX * 4, "d2h_bytes_per_step": NVOX * 4 + 4 + 32 * max(nboxes[0], 1)},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
        if latency is not None:
            line["latency"] = latency
        if sweep is not None:
            line["sweep"] = sweep
        if train is not None:
            line["train"] = train
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if eager is not None:
            line["eager_gpu"] = eager
        emit(line)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


def window_batch():
    import light_unet.utils as lu
    return lu.WINDOW_BATCH


def eager_gpu_leg(variant, device):
    """Informational: the reference's own path on THIS GPU -- the oracle's restatement of utils.py:86-137 / unet3d.py run
    with stock ATen / cuDNN kernels (`.cuda()`, fp32, TF32 off), batch-1 forward per window as the reference does, on a
    bounded sample (65 windows).  SURVEY 0.1: PyTorch eager is the only existing GPU implementation."""
    from oracle import synth, unet_ref
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg, sd = cpu_reference_setup(variant)
    sd = {k: v.to(device) for k, v in sd.items()}
    vol = torch.from_numpy(synth.synth_volume(VOLUME, seed=42, n_blobs=6)).to(device)
    wins = [vol[0:48, y:y + 48, x:x + 48] for y in (0, 24, 48, 72, 80) for x in list(range(0, 265, 24)) + [272]]   # 65 windows

    def run(batch):
        with torch.no_grad():
            for i in range(0, len(wins), batch):
                unet_ref.forward(sd, torch.stack(wins[i:i + batch])[:, None], cfg)
    out = {}
    for batch in (1, 13):
        run(batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(batch); e1.record()
        torch.cuda.synchronize()
        wps = len(wins) / (e0.elapsed_time(e1) * 1e-3)
        out[f"batch{batch}"] = {"windows_per_s": round(wps, 1), "voxels_per_s": round(wps * NVOX / NWIN, 1)}
    out["note"] = ("oracle port on cuda (ATen/cuDNN eager, fp32, TF32 off), forward only, 65 of the 325 windows; batch1 is how "
                   "the reference runs (utils.py:115-118)")
    return out


def bench_train(args, device, rank, world, dtype):
    """configs[1]/[3]: training step on 48^3 patches, batch 8 per GPU, data parallel over NCCL when world > 1."""
    from light_unet.models import Lightweight3DUNet, FocalTverskyLoss
    from light_unet.engine import UNetPlan
    if not hasattr(UNetPlan, "backward"):
        return {"unavailable": "backward kernels not built in this revision"}
    from light_unet.parallel import DataParallelStep
    from oracle import synth, unet_ref
    kw = VARIANTS[args.variant]
    cfg = unet_ref.UNetCfg(dropout_p=0.1, **kw)
    model = Lightweight3DUNet(dropout_p=0.1, **kw)
    model.load_state_dict(unet_ref.to_torch(synth.synth_state_dict(unet_ref.param_shapes(cfg), 1)))
    model = model.to(device).set_compute_dtype(dtype).train()
    graph = args.variant == "dws" and not args.no_graph
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-5, fused=True, capturable=graph)     # trainer.py:75-79
    stepper = DataParallelStep(model, FocalTverskyLoss(), opt, world_size=world, use_graph=graph)
    B = TRAIN_BATCH
    hx, ht = [], []
    for i in range(4):
        x, t = synth.synth_patches(B, 48, seed=100 + 17 * rank + i)
        hx.append(torch.from_numpy(x).pin_memory())
        ht.append(torch.from_numpy(t).pin_memory())
    dx, dtg = [h.to(device) for h in hx], [h.to(device) for h in ht]
    last = [None]

    def step_resident(i):
        last[0] = stepper.step(dx[i % 4], dtg[i % 4])

    def step_e2e(i):
        # the batch of step i was staged (pinned host -> device on a side stream) while step i - 1 ran; the loss is read
        # every step as trainer.py:234 does
        loss = stepper.step()
        stepper.prefetch(hx[(i + 1) % 4], ht[(i + 1) % 4])
        last[0] = loss.item()

    steps = max(args.steps * 4, 10)
    for i in range(max(args.warmup, 3)):
        step_resident(i)
    ms = timed(step_resident, steps, world, device)
    stepper.prefetch(hx[0], ht[0])
    for i in range(2):
        step_e2e(i)
    ms_e2e = timed(step_e2e, steps, world, device)
    stepper._staged = None
    # the whole training pipeline on the device: DevicePatchSampler (class-balanced crop + the FL-70 augmentations,
    # configs/unet_fl70.yaml:12-50, from volumes resident in HBM) feeding the graph step
    pipeline = None
    if not args.quick:
        from light_unet.datasets import DevicePatchSampler
        aug = {"random_flip": {"enabled": True, "prob": 0.5, "axes": [0, 1, 2]},
               "random_rotation": {"enabled": True, "prob": 0.5, "angle_range": [-15, 15], "axes": [[0, 1], [0, 2], [1, 2]]},
               "random_scale": {"enabled": True, "prob": 0.3, "scale_range": [0.9, 1.1]},
               "intensity_shift": {"enabled": True, "prob": 0.5, "shift_range": [-0.1, 0.1]},
               "gaussian_noise": {"enabled": True, "prob": 0.3, "sigma": 0.01}}
        cases = []
        for i in range(4):
            v = synth.synth_volume(VOLUME, seed=300 + 7 * rank + i, n_blobs=8)
            cases.append((v, (v > 0.55).astype(np.float32)))
        sampler = DevicePatchSampler(cases, PATCH, 0.5, aug, seed=42 + rank, device=device)

        def step_pipeline(i):
            x, t = sampler.sample_batch(B)
            last[0] = stepper.step(x, t)
        for i in range(3):
            step_pipeline(i)
        ms_pipe = timed(step_pipeline, steps, world, device)
        pipeline = {"value": round(world * B * steps / (ms_pipe * 1e-3), 1), "unit": "patches/s", "ms_per_step": round(ms_pipe / steps, 3),
                    "what": "DevicePatchSampler.sample_batch (4 cached 128x128x320 cases, FL-70 augmentations) + the training step"}
    pk = peaks()
    es = 2 if dtype == "f16" else 4
    per_step_s = ms * 1e-3 / steps
    parity = ("every parameter gradient as close to the float64 oracle as the reference's own fp32 arithmetic (tests/test_gpu_configs.py)"
              if dtype == "f32" else "loss 1e-4, probabilities 1e-2, whole-gradient cosine > 0.98 (per-tensor parity not claimed)")
    return {"metric": "train 48^3 patches/s", "value": round(world * B * steps / (ms * 1e-3), 1), "unit": "patches/s",
            "dtype": dtype, "parity": parity,
            "ms_per_step": round(ms / steps, 3), "steps": steps, "global_batch": world * B,
            "workload": "configs[1]: fwd + FocalTversky + bwd + AdamW(lr 1e-4, wd 1e-5), dropout 0.1, batch 8 of 48^3 per GPU",
            "parallelism": f"dp{world}", "cuda_graph": bool(stepper.use_graph and stepper._graph is not None),
            "loss": float(last[0]) if last[0] is not None else None,
            "e2e": {"value": round(world * B * steps / (ms_e2e * 1e-3), 1), "unit": "patches/s",
                    "h2d_bytes_per_step": 2 * B * 48 ** 3 * 4, "d2h_bytes_per_step": 4},
            "with_device_sampler": pipeline,
            # forward compulsory traffic (write + read) x3 (fwd + ~2x for bwd)
            "frac_hbm": round(3 * B * ELEMS_PER_PATCH * es * 2 / per_step_s / 1e9 / pk["hbm_gbs"], 4)}


# ----------------------------------------------------------------------------------- CPU reference arm
def cpu_reference_setup(variant):
    from oracle import synth, unet_ref
    cfg = unet_ref.UNetCfg(dropout_p=0.0, **VARIANTS[variant])
    sd = unet_ref.to_torch(synth.synth_state_dict(unet_ref.param_shapes(cfg), 3))
    return cfg, sd


def cpu_sample(cfg, sd, vol):
    """The reference algorithm (utils.py:86-137, one batch-1 forward per window; inferencer.py:62-111) on a
    sub-volume, through the oracle's restatement (torch CPU ATen kernels, all host threads)."""
    from oracle import bbox_ref, stitch_ref, unet_ref

    def predict(chunk):
        with torch.no_grad():
            return unet_ref.forward(sd, torch.from_numpy(chunk), cfg).numpy()
    prob = stitch_ref.sliding_window(vol, predict, PATCH, 0.5, True, batch=1)
    boxes = bbox_ref.extract_bboxes(prob, 0.3, 0.5, (4.0, 4.0, 4.0), 3)
    return prob, boxes


def cpu_baseline(variant, target_s=12.0):
    from oracle import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, sd = cpu_reference_setup(variant)
    full = synth.synth_volume(VOLUME, seed=42, n_blobs=6)
    small = np.ascontiguousarray(full[:48, :48, :96])          # 1 x 1 x 3 windows: warm-up + rate estimate
    cpu_sample(cfg, sd, small)
    t0 = time.perf_counter()
    cpu_sample(cfg, sd, small)
    per_win = (time.perf_counter() - t0) / 3
    # bounded sample: a z-slab of the same volume holding ~target_s of work (full y/x extent when affordable)
    nwin_target = max(3, int(target_s / per_win))
    if nwin_target >= NWIN:
        sub, nwin = full, NWIN                                   # the whole volume
    elif nwin_target >= 65:
        nz = min(4, nwin_target // 65)
        sub, nwin = full[:48 + 24 * (nz - 1)], nz * 65           # nz x 5 x 13 windows
    elif nwin_target >= 13:
        ny = min(5, nwin_target // 13)
        ext = 48 + 24 * (ny - 1) if ny < 5 else 128
        sub, nwin = full[:48, :ext], ny * 13
    else:
        nx = max(3, nwin_target)
        sub, nwin = full[:48, :48, :48 + 24 * (nx - 1)], nx
    sub = np.ascontiguousarray(sub)
    t0 = time.perf_counter()
    cpu_sample(cfg, sd, sub)
    dt = time.perf_counter() - t0
    vox_per_win = NVOX / NWIN                                    # the full volume costs 325 windows for 5.24 M voxels
    return {"value": round(nwin / dt * vox_per_win, 1), "unit": "voxels/s", "cores": cores, "kind": "port",
            "windows_per_s": round(nwin / dt, 2),
            "sample": f"{nwin} of the 325 windows (sub-volume {tuple(sub.shape)} of the same synthetic volume), batch-1 "
                      f"forward per window + stitch + bbox, torch CPU fp32 with {cores} threads, {dt:.1f} s; "
                      f"voxels/s = windows/s x {vox_per_win:.0f} volume-voxels per window"}


def bench_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    from oracle import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg, sd = cpu_reference_setup(args.variant)
    full = synth.synth_volume(VOLUME, seed=42, n_blobs=6)
    sub = np.ascontiguousarray(full[:48])                       # one z-slab of the window grid: 1 x 5 x 13 = 65 windows per step
    nwin = 65
    for _ in range(args.warmup):
        cpu_sample(cfg, sd, sub)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_sample(cfg, sd, sub)
    dt = time.perf_counter() - t0
    vox_per_win = NVOX / NWIN
    v = round(args.steps * nwin / dt * vox_per_win, 1)
    sample = (f"each step = {nwin} of the 325 windows (sub-volume {tuple(sub.shape)}), batch-1 forward per window + stitch + "
              f"bbox through the oracle port of the reference CPU path, torch CPU fp32, {cores} threads; voxels/s = "
              f"windows/s x {vox_per_win:.0f}")
    line = {"impl": "reference", "metric": "sliding-window inference volume-voxels/s", "value": v, "unit": "voxels/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * dt / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[2]: sliding-window inference, synthetic whole-body PET 128x128x320 @4mm, "
                                   "48^3 windows, 50% overlap (325 windows), Gaussian stitch, threshold 0.3 -> CC -> bbox",
                       "variant": args.variant, "sample": sample},
            "cpu_baseline": {"value": v, "unit": "voxels/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_JSON_FD = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's original stdout; everything else (NCCL's version banner, library
    chatter) was redirected to stderr at start-up so that stdout holds nothing but this line."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_JSON_FD if _JSON_FD is not None else 1, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="dws", choices=sorted(VARIANTS))
    ap.add_argument("--dtype", default="f16", choices=["f16", "f32"])
    ap.add_argument("--train-dtype", default="f32", choices=["f16", "f32"])
    ap.add_argument("--sweep", action="store_true", help="configs[4]: patch 48/64/96 x batch 1..64 forward throughput")
    ap.add_argument("--patch", type=int, default=0, help="with --batch: one sweep point as the headline workload")
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--eager-gpu", action="store_true", help="informational: the oracle port on cuda (ATen/cuDNN eager)")
    ap.add_argument("--no-latency", action="store_true", help="N > 1: skip the window-sharded single-volume latency leg")
    ap.add_argument("--quick", action="store_true", help="skip the secondary legs (fp16-storage training)")
    ap.add_argument("--no-graph", action="store_true", help="training step launched eagerly instead of as one CUDA graph")
    ap.add_argument("--skip-train", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        bench_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback); use --impl reference for the CPU arm")
    _, world, _ = dist_env()
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            raise SystemExit("bench.py: launch N>1 with torchrun (python -m torch.distributed.run --nproc-per-node N ...)")
    bench_ours(args)


if __name__ == "__main__":
    main()

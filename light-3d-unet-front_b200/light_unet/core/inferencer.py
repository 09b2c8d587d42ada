"""Inferencer -- drop-in for light_unet/core/inferencer.py.

Same constructor, `extract_bboxes`, `infer_case` and `infer_split` contract as the
reference; the sliding window, threshold, connected components and per-component
reductions run on the GPU (libl3d).  `infer_volume` is the device-resident core
(volume in -> probability map + candidate boxes out) that `infer_case` wraps with
the NIfTI / JSON file I/O of the reference (inferencer.py:122-134,164-180; needs
nibabel, imported lazily because file I/O is off the accelerated path).
"""
from __future__ import annotations

import json
import os
from pathlib import Path

import numpy as np
import torch

from .. import _native as nv
from ..models.metrics import label_device
from ..models.unet3d import Lightweight3DUNet
from ..utils import find_case_files, sliding_window_device, sliding_window_inference_3d
from .config import ConfigManager

BBOX_TABLE_CAP = 1 << 16


def _load_nifti(path):
    """-> (float32 array in nibabel index order, affine, header with get_zooms()): nibabel as in the reference
    (inferencer.py:122-125) when it is installed, the built-in NIfTI-1 reader (io_nifti) otherwise."""
    try:
        import nibabel as nib
    except ImportError:
        from .. import io_nifti
        data, hdr = io_nifti.load(path)
        return data, hdr.affine, hdr
    img = nib.load(path)
    return img.get_fdata().astype(np.float32), img.affine, img.header


def _save_nifti(data, affine, header, path):
    """inferencer.py:164-165."""
    try:
        import nibabel as nib
    except ImportError:
        from .. import io_nifti
        io_nifti.save(data, affine, header if isinstance(header, io_nifti.NiftiHeader) else None, path)
        return
    nib.save(nib.Nifti1Image(data, affine, header), path)


class Inferencer:
    """Inference class for generating predictions (reference: inferencer.py:18-60)."""

    def __init__(self, config_or_path, model_path):
        if isinstance(config_or_path, (str, Path)):
            self.config = ConfigManager.load(str(config_or_path))
        elif isinstance(config_or_path, dict):
            self.config = config_or_path
        else:
            raise TypeError(f"config_or_path must be str, Path or dict, got {type(config_or_path)}")
        if not torch.cuda.is_available():
            raise nv.NativeError("Inferencer: the B200-native path needs a CUDA device (no CPU fallback)")
        self.device = torch.device("cuda", torch.cuda.current_device())
        self._copy_stream = None
        print(f"Using device: {self.device}")
        m = self.config["model"]
        self.model = Lightweight3DUNet(
            in_channels=1, out_channels=m["output_channels"], start_channels=m["start_channels"],
            encoder_channels=m["encoder_channels"], use_depthwise_separable=m["use_depthwise_separable"],
            use_grouped=m["use_grouped_conv"], groups=m["groups"], dropout_p=0.0).to(self.device)
        checkpoint = torch.load(model_path, map_location=self.device)
        self.model.load_state_dict(checkpoint["model_state_dict"])
        self.model.eval()
        print(f"Loaded model from {model_path}")
        print(f"Best epoch: {checkpoint.get('best_epoch', 'N/A')}")
        print(f"Best metric: {checkpoint.get('best_metric', 'N/A'):.4f}")
        self.prob_maps_dir = Path(self.config["output"]["prob_maps_dir"])
        self.bboxes_dir = Path(self.config["output"]["bboxes_dir"])
        self.prob_maps_dir.mkdir(parents=True, exist_ok=True)
        self.bboxes_dir.mkdir(parents=True, exist_ok=True)

    # ------------------------------------------------------------------ boxes
    @staticmethod
    def _min_voxels(min_volume_cc, spacing):
        voxel_volume_cc = spacing[0] * spacing[1] * spacing[2] / 1000.0
        return int(np.ceil(min_volume_cc / voxel_volume_cc))                      # inferencer.py:66-68

    def _label_and_reduce(self, prob_d: torch.Tensor, mask_d: torch.Tensor, min_voxels: int):
        """Device part of extract_bboxes: mask -> labels -> per-component table (no host synchronisation)."""
        D, H, W = prob_d.shape
        labels, n_d = label_device(mask_d, min_voxels if min_voxels > 0 else 0)
        st = nv.stream_ptr(prob_d.device)
        table = torch.empty(BBOX_TABLE_CAP, 8, dtype=torch.int32, device=prob_d.device)
        nv.call("l3d_bbox_init", nv.ptr(table), BBOX_TABLE_CAP, st)
        nv.call("l3d_bbox_reduce", nv.ptr(labels), nv.ptr(prob_d), D, H, W, nv.ptr(table), BBOX_TABLE_CAP, st,
                algo_bytes=8 * D * H * W)                          # labels + probabilities read once
        return labels, n_d, table

    def _bboxes_from_device(self, prob_d: torch.Tensor, mask_d: torch.Tensor, min_volume_cc, spacing, reduced=None):
        """threshold mask -> labels -> per-component table on the device, then the reference's host-side
        arithmetic on the resulting integers (inferencer.py:66-108)."""
        D, H, W = prob_d.shape
        voxel_volume_cc = spacing[0] * spacing[1] * spacing[2] / 1000.0
        min_voxels = self._min_voxels(min_volume_cc, spacing)
        labels, n_d, table = reduced if reduced is not None else self._label_and_reduce(prob_d, mask_d, min_voxels)
        st = nv.stream_ptr(prob_d.device)
        n = int(n_d.item())
        if n > BBOX_TABLE_CAP:
            cap = n
            table = torch.empty(cap, 8, dtype=torch.int32, device=prob_d.device)
            nv.call("l3d_bbox_init", nv.ptr(table), cap, st)
            nv.call("l3d_bbox_reduce", nv.ptr(labels), nv.ptr(prob_d), D, H, W, nv.ptr(table), cap, st)
        return self._boxes_from_rows(table[:n].cpu().numpy(), (D, H, W), spacing)

    def _label_table(self, labels, prob_d, n):
        """A box table large enough for n components from an existing label map."""
        D, H, W = prob_d.shape
        st = nv.stream_ptr(prob_d.device)
        table = torch.empty(max(n, 1), 8, dtype=torch.int32, device=prob_d.device)
        nv.call("l3d_bbox_init", nv.ptr(table), max(n, 1), st)
        nv.call("l3d_bbox_reduce", nv.ptr(labels), nv.ptr(prob_d), D, H, W, nv.ptr(table), max(n, 1), st)
        return labels, torch.tensor([n], dtype=torch.int32, device=prob_d.device), table

    def _boxes_from_rows(self, rows, shape, spacing):
        """The reference's host-side arithmetic on the per-component integers (inferencer.py:83-108)."""
        voxel_volume_cc = spacing[0] * spacing[1] * spacing[2] / 1000.0
        expansion = self.config["data"]["bbox_expansion_voxels"]
        out = []
        for i in range(len(rows)):
            row = rows[i]
            if row[6] == 0:
                continue
            lo = [np.int64(row[0]), np.int64(row[2]), np.int64(row[4])]
            hi = [np.int64(row[1]), np.int64(row[3]), np.int64(row[5])]
            box = []
            for ax in range(3):
                box.append(max(0, lo[ax] - expansion))
                box.append(min(shape[ax] - 1, hi[ax] + expansion))
            mm = [box[2 * ax + k] * spacing[ax] for ax in range(3) for k in range(2)]
            volume_cc = np.int64(row[6]) * voxel_volume_cc
            confidence = np.array([row[7]], dtype=np.int32).view(np.float32)[0]
            out.append({"mask_id": i + 1,
                        "bbox_voxel": [int(v) for v in box],
                        "bbox_mm": [float(v) for v in mm],
                        "volume_cc": float(volume_cc),
                        "confidence": float(confidence)})
        return out

    def extract_bboxes(self, prob_map, threshold=0.3, min_volume_cc=0.5, spacing=(4.0, 4.0, 4.0)):
        """Extract bounding boxes from a probability map (reference signature, inferencer.py:62)."""
        if isinstance(prob_map, torch.Tensor):
            prob_d = prob_map.to(self.device, dtype=torch.float32).contiguous()
        else:
            prob_d = torch.from_numpy(np.ascontiguousarray(prob_map, dtype=np.float32)).to(self.device)
        if prob_d.dim() != 3:
            raise ValueError(f"Expected 3D probability map, got shape {tuple(prob_d.shape)}")
        mask_d = torch.empty(prob_d.shape, dtype=torch.int32, device=self.device)
        # `prob_map >= threshold` compares float32 values with the python float converted to float32
        nv.call("l3d_threshold", nv.ptr(prob_d), prob_d.numel(), float(np.float32(threshold)), nv.ptr(mask_d),
                nv.stream_ptr(self.device))
        return self._bboxes_from_device(prob_d, mask_d, min_volume_cc, spacing)

    # ----------------------------------------------------------------- volume
    def infer_volume(self, image, threshold=0.3, spacing=(4.0, 4.0, 4.0), body_mask=None, return_device=False, prob_out=None,
                     shard=None):
        """Device-resident case pipeline: sliding window -> (body mask) -> threshold -> CC -> boxes.
        `image` is a host ndarray or CUDA tensor [D,H,W].  Returns (prob_map, bboxes).  `prob_out`: optional pinned host
        fp32 tensor [D,H,W]; the probability map is then copied into it on a side stream while the connected-component /
        bounding-box kernels run, and returned instead of a fresh array.
        `shard = (rank, world_size, group)`: the windows of THIS volume are split over the ranks of a torch.distributed
        group (every rank calls with the same volume; parallel/window_shard.py); rank 0 returns the result, the other
        ranks (None, [])."""
        if isinstance(image, torch.Tensor):
            vol = image.to(self.device, dtype=torch.float32, non_blocking=True)
        else:
            vol = torch.from_numpy(np.ascontiguousarray(image, dtype=np.float32)).to(self.device, non_blocking=True)
        bm = None
        if body_mask is not None:
            bm = body_mask if isinstance(body_mask, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(body_mask).astype(np.uint8))
        min_cc = self.config["data"]["volume_threshold"]["inference_cc"]
        reduced = None
        if shard is None and not nv.TIMER.enabled and os.environ.get("L3D_INFER_GRAPH", "1") != "0":
            prob_d, mask_d, reduced = self._graph_pipeline(vol, bm, threshold, self._min_voxels(min_cc, spacing))
            if return_device:
                prob_d = prob_d.clone()                           # the graph's output buffer is overwritten by the next call
        else:
            prob_d, mask_d = sliding_window_device(vol, self.model, tuple(self.config["data"]["patch_size"]), 0.5, True,
                                                   body_mask=bm, threshold=threshold, shard=shard)
        if prob_d is None:                                        # window-sharded call on a rank other than 0
            return None, []
        copy_done = None
        if prob_out is not None:
            if not (isinstance(prob_out, torch.Tensor) and prob_out.device.type == "cpu" and prob_out.dtype == torch.float32
                    and tuple(prob_out.shape) == tuple(prob_d.shape) and prob_out.is_contiguous()):
                raise ValueError("prob_out must be a contiguous float32 CPU tensor of the volume's shape")
            if getattr(self, "_copy_stream", None) is None:
                self._copy_stream = torch.cuda.Stream(device=self.device)
            main = torch.cuda.current_stream(self.device)
            self._copy_stream.wait_stream(main)                   # the stitched map is complete on the main stream
            with torch.cuda.stream(self._copy_stream):
                prob_out.copy_(prob_d, non_blocking=True)
                prob_d.record_stream(self._copy_stream)
                copy_done = torch.cuda.Event()
                copy_done.record(self._copy_stream)
        bboxes = self._bboxes_from_device(prob_d, mask_d, min_cc, spacing, reduced)
        if copy_done is not None:
            copy_done.synchronize()
            return prob_out, bboxes
        return (prob_d if return_device else prob_d.cpu().numpy()), bboxes

    def infer_volumes(self, images, threshold=0.3, spacing=(4.0, 4.0, 4.0), prob_outs=None):
        """Generator over a sequence of host volumes (pinned tensors copy asynchronously): yields (prob_map, bboxes) per
        volume like infer_volume, in order, as a three-stage pipeline -- the host-to-device copy of volume i + 1 and the
        device-to-host copy of map i - 1 (+ its box table) run on side streams under the kernels of volume i, so the result
        of a volume is yielded once the next one has been launched.  `prob_outs`: optional sequence of pinned host fp32
        tensors receiving the maps (fresh pinned tensors otherwise)."""
        dev = self.device
        if getattr(self, "_h2d_stream", None) is None:
            self._h2d_stream = torch.cuda.Stream(device=dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        it = iter(images)
        outs = iter(prob_outs) if prob_outs is not None else None
        min_cc = self.config["data"]["volume_threshold"]["inference_cc"]
        min_voxels = self._min_voxels(min_cc, spacing)
        rows_cap = 4096                                            # box-table rows fetched with the map; more -> second fetch
        slots = [None, None]

        def stage(img):
            host = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(img, dtype=np.float32))
            with torch.cuda.stream(self._h2d_stream):
                d = host.to(dev, dtype=torch.float32, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._h2d_stream)
            return d, ev

        def launch(vol, ev, k):
            cur = torch.cuda.current_stream(dev)
            cur.wait_event(ev)
            vol.record_stream(cur)
            if nv.TIMER.enabled or os.environ.get("L3D_INFER_GRAPH", "1") == "0":
                prob_d, mask_d = sliding_window_device(vol, self.model, tuple(self.config["data"]["patch_size"]), 0.5, True, threshold=threshold)
                labels, n_d, table = self._label_and_reduce(prob_d, mask_d, min_voxels)
            else:
                prob_d, mask_d, (labels, n_d, table) = self._graph_pipeline(vol, None, threshold, min_voxels)
            sl = slots[k & 1]
            if sl is None or sl["prob"].shape != prob_d.shape:
                sl = slots[k & 1] = {"prob": torch.empty_like(prob_d), "labels": torch.empty_like(labels),
                                     "n": torch.empty(1, dtype=torch.int32).pin_memory(), "rows": torch.empty(rows_cap, 8, dtype=torch.int32).pin_memory()}
            # the graph's output buffers are overwritten by the next volume: keep device copies (21 MB each, microseconds)
            sl["prob"].copy_(prob_d, non_blocking=True)
            sl["labels"].copy_(labels, non_blocking=True)
            sl["n"].copy_(n_d, non_blocking=True)
            sl["rows"].copy_(table[:rows_cap], non_blocking=True)
            done = torch.cuda.Event()
            done.record(cur)
            host_out = next(outs) if outs is not None else torch.empty(tuple(prob_d.shape), dtype=torch.float32).pin_memory()
            self._copy_stream.wait_event(done)
            with torch.cuda.stream(self._copy_stream):
                host_out.copy_(sl["prob"], non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(self._copy_stream)
            return sl, done, copied, host_out

        def finish(job):
            sl, done, copied, host_out = job
            done.synchronize()
            n = int(sl["n"][0])
            if n > rows_cap:                                       # rare: more components than fetched rows
                boxes = self._bboxes_from_device(sl["prob"], None, min_cc, spacing, reduced=self._label_table(sl["labels"], sl["prob"], n))
            else:
                boxes = self._boxes_from_rows(sl["rows"][:n].numpy(), tuple(sl["prob"].shape), spacing)
            copied.synchronize()
            return host_out, boxes

        try:
            nxt = stage(next(it))
        except StopIteration:
            return
        pending, k = None, 0
        while nxt is not None:
            vol, ev = nxt
            try:
                nxt = stage(next(it))                              # upload of the next volume overlaps this one's kernels
            except StopIteration:
                nxt = None
            job = launch(vol, ev, k)
            k += 1
            if pending is not None:
                yield finish(pending)                              # the previous volume's map / boxes, copied under this one
            pending = job
        yield finish(pending)

    # ------------------------------------------------------------- CUDA graph
    def _graph_pipeline(self, vol, bm, threshold, min_voxels):
        """Window gather -> network -> stitch / threshold -> labelling -> box table as ONE CUDA graph per (volume shape,
        threshold, patch, storage type): the ~45 launches of a case cost one host call.  The first call of a configuration
        runs eagerly (it sizes the workspace and fills the caches the capture must not touch), the second captures, later
        ones replay.  A graph is dropped when the model's activation workspace was replaced in the meantime."""
        patch = tuple(self.config["data"]["patch_size"])
        plan = self.model._plan
        key = (tuple(vol.shape), float(np.float32(threshold)), bm is not None, patch, self.model.compute_dtype, int(min_voxels), str(vol.device))
        graphs = self.__dict__.setdefault("_graphs", {})
        seen = self.__dict__.setdefault("_graph_seen", set())
        g = graphs.get(key)
        if g is not None and plan._ws.get((patch, self.model.compute_dtype, str(vol.device), False)) is not g["ws"]:
            graphs.pop(key)
            g = None
        if g is None:
            def run(v, m):
                prob_d, mask_d = sliding_window_device(v, self.model, patch, 0.5, True, body_mask=m, threshold=threshold)
                return prob_d, mask_d, self._label_and_reduce(prob_d, mask_d, min_voxels)
            if key not in seen:                                   # first call: eager
                seen.add(key)
                return run(vol, bm)
            if len(graphs) >= 4:
                graphs.pop(next(iter(graphs)))
            sv = vol.clone()
            sm = bm.to(device=vol.device, dtype=torch.uint8).clone() if bm is not None else None
            graph = torch.cuda.CUDAGraph()
            torch.cuda.synchronize(vol.device)
            l0 = nv.launch_count()
            with torch.cuda.graph(graph):
                outs = run(sv, sm)
            g = {"graph": graph, "vol": sv, "bm": sm, "outs": outs, "launches": nv.launch_count() - l0,
                 "ws": plan._ws.get((patch, self.model.compute_dtype, str(vol.device), False))}
            graphs[key] = g
        g["vol"].copy_(vol, non_blocking=True)
        if bm is not None:
            g["bm"].copy_(bm, non_blocking=True)
        g["graph"].replay()
        nv.add_launch_count(g["launches"])
        return g["outs"]

    # ------------------------------------------------------------- file level
    def infer_case(self, case_id, data_dir, threshold=0.3):
        """Single-case inference with the reference's file layout (inferencer.py:113-183)."""
        data_dir = Path(data_dir)
        image_files = find_case_files(data_dir, case_id, file_type="image")
        if len(image_files) == 0:
            print(f"Warning: No image files found for {case_id}")
            return False
        image, affine, header = _load_nifti(image_files[0])
        spacing = [float(s) for s in header.get_zooms()[:3]]
        bm_cfg = self.config.get("data", {}).get("body_mask", {})
        apply_bm = bm_cfg.get("apply_to_inference", False) and bm_cfg.get("enabled", False)
        body_mask = None
        if apply_bm:
            bm_path = data_dir / "body_masks" / f"{case_id}.nii.gz"
            if bm_path.exists():
                body_mask = _load_nifti(bm_path)[0].astype(bool)
            else:
                print(f"Warning: Body mask not found for {case_id}")
        print(f"Running inference on {case_id}...")
        try:
            prob_map, bboxes = self.infer_volume(image, threshold=threshold, spacing=spacing, body_mask=body_mask)
        except Exception as e:   # the reference swallows sliding-window failures and reports the case as failed
            print(f"Error during inference execution for {case_id}: {e}")
            return False
        _save_nifti(prob_map, affine, header, self.prob_maps_dir / f"{case_id}_prob.nii.gz")
        bbox_json = {"case_id": case_id, "processing_path": "B", "orig_spacing": spacing, "threshold": threshold,
                     "num_candidates": len(bboxes), "candidates": bboxes}
        bbox_path = self.bboxes_dir / f"{case_id}_bboxes.json"
        with open(bbox_path, "w") as f:
            json.dump(bbox_json, f, indent=2)
        print(f"Found {len(bboxes)} candidates, saved to {bbox_path}")
        return True

    def infer_split(self, split_file, data_dir):
        """All cases of a split file (inferencer.py:185-201)."""
        with open(split_file, "r") as f:
            case_ids = [line.strip() for line in f if line.strip()]
        # one process per GPU (torch.distributed initialised): cases are independent, so case i goes to rank i mod N and no
        # data-path collective is needed (SURVEY.md 8(e) partition (1)); only the summary counts are reduced
        import torch.distributed as dist
        from ..parallel import shard_cases
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        rank = dist.get_rank() if world > 1 else 0
        mine = shard_cases(case_ids, rank, world)
        print(f"Performing inference on {len(case_ids)} cases..." + (f" (rank {rank}/{world}: {len(mine)})" if world > 1 else ""))
        successful, failed = 0, []
        threshold = self.config["validation"]["default_threshold"]
        for case_id in mine:
            if self.infer_case(case_id, data_dir, threshold=threshold):
                successful += 1
            else:
                failed.append(case_id)
        if world > 1:
            t = torch.tensor([successful], dtype=torch.int64, device=self.device if dist.get_backend() == "nccl" else "cpu")
            dist.all_reduce(t)
            successful = int(t.item())
            gathered = [None] * world
            dist.all_gather_object(gathered, failed)
            failed = [c for part in gathered for c in part]
        print(f"\nInference complete: Successful: {successful}/{len(case_ids)}")

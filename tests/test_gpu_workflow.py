"""System test of the whole training workflow on the device (SURVEY.md 8(a16), 8(f) N1-N3 together), on synthetic cases:
DevicePatchSampler -> DataParallelStep (one CUDA graph: forward, Focal Tversky, backward, AdamW) for a few dozen steps ->
DeviceValidator sweep -> checkpoint in the reference's format -> Inferencer.  Checks that the pieces compose: the loss goes
down on a learnable toy task, validation returns the reference's metric keys, the checkpoint round-trips."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cases(n, shape, seed):
    """Bright blobs on a dim noisy background; the label is the blob mask -- learnable from intensity alone."""
    rng = np.random.default_rng(seed)
    out = []
    zz, yy, xx = np.meshgrid(*[np.arange(s, dtype=np.float32) for s in shape], indexing="ij")
    for _ in range(n):
        img = (0.15 * rng.random(shape, dtype=np.float32)).astype(np.float32)
        lab = np.zeros(shape, dtype=np.float32)
        for _ in range(int(rng.integers(3, 6))):
            c = [rng.uniform(4, s - 5) for s in shape]
            r = rng.uniform(2.5, 4.5)
            m = ((zz - c[0]) ** 2 + (yy - c[1]) ** 2 + (xx - c[2]) ** 2) <= r * r
            img[m] = np.clip(0.75 + 0.2 * rng.random(int(m.sum()), dtype=np.float32), 0, 1)
            lab[m] = 1.0
        out.append((img, lab))
    return out


def test_train_validate_checkpoint_infer(tmp_path):
    from light_unet.core.inferencer import Inferencer
    from light_unet.core.validation import DeviceValidator, use_device_validation
    from light_unet.datasets import DevicePatchSampler
    from light_unet.models import Lightweight3DUNet, get_loss_function
    from light_unet.parallel import DataParallelStep
    torch.manual_seed(0)
    train_cases, val_cases = _cases(3, (40, 40, 48), 1), _cases(2, (32, 40, 40), 2)
    aug = {"random_flip": {"enabled": True, "prob": 0.5, "axes": [0, 1, 2]}, "intensity_shift": {"enabled": True, "prob": 0.5, "shift_range": [-0.05, 0.05]}}
    sampler = DevicePatchSampler(train_cases, (16, 16, 16), 0.7, aug, seed=42, device=torch.device(DEV))
    model = Lightweight3DUNet(dropout_p=0.1).to(DEV).set_compute_dtype("f32").train()
    loss_fn = get_loss_function({"name": "FocalTverskyLoss", "alpha": 0.7, "beta": 0.3, "gamma": 0.75})
    opt = torch.optim.AdamW(model.parameters(), lr=3e-3, weight_decay=1e-5, fused=True, capturable=True)
    stepper = DataParallelStep(model, loss_fn, opt, use_graph=True)
    losses = []
    for it in range(120):
        x, t = sampler.sample_batch(8)
        losses.append(float(stepper.step(x, t)))
    first, last = float(np.mean(losses[:10])), float(np.mean(losses[-10:]))
    print(f"loss {first:.4f} -> {last:.4f} over 120 graph-replayed steps of 8 x 16^3 patches")
    assert stepper._graph is not None and np.isfinite(losses).all() and last < 0.8 * first
    # ---- validation on the device, through the reference's loader contract and through the Trainer hook
    config = {"validation": {"default_threshold": 0.5, "threshold_sensitivity_range": [0.3, 0.5, 0.7]},
              "metrics": {"model_selection": {"tie_threshold": 0.01}},
              "data": {"patch_size": [16, 16, 16], "spacing": {"target": [4.0, 4.0, 4.0]}, "bbox_expansion_voxels": 2,
                       "volume_threshold": {"inference_cc": 0.1}},
              "model": {"output_channels": 1, "start_channels": 16, "encoder_channels": [16, 32, 64, 128], "use_depthwise_separable": True,
                        "use_grouped_conv": True, "groups": 8},
              "output": {"prob_maps_dir": str(tmp_path / "prob"), "bboxes_dir": str(tmp_path / "bbox")}}
    loader = [(torch.from_numpy(img)[None, None], torch.from_numpy(lab)[None, None], [f"{i:04d}"], torch.tensor([[4.0, 4.0, 4.0]]))
              for i, (img, lab) in enumerate(val_cases)]
    zero, metrics = DeviceValidator(model, config).validate(loader)
    for k in ("lesion_wise_recall", "lesion_wise_precision", "lesion_wise_f1", "voxel_wise_dsc_micro", "voxel_wise_dsc_macro", "fp_per_case", "tp", "fp",
              "fn", "best_threshold", "best_recall", "best_dsc_macro", "dsc", "recall", "precision"):
        assert k in metrics, k
    print(f"validation: best threshold {metrics['best_threshold']}, lesion recall {metrics['best_recall']:.2f}, macro Dice {metrics['best_dsc_macro']:.3f}")
    assert zero == 0.0 and metrics["best_dsc_macro"] > 0.3 and not model.training

    class FakeTrainer:          # the attributes the reference's Trainer.validate reads (trainer.py:349-445)
        pass
    tr = FakeTrainer()
    tr.model, tr.config, tr.val_loader, tr.device = model, config, loader, torch.device(DEV)
    use_device_validation(tr)
    assert tr.validate(0) == (zero, metrics)
    # ---- checkpoint in the reference's format (trainer.py:448-459) -> Inferencer -> boxes
    ckpt = tmp_path / "best_model.pth"
    torch.save({"epoch": 0, "model_state_dict": model.state_dict(), "best_epoch": 0, "best_metric": float(metrics["best_recall"])}, ckpt)
    inf = Inferencer(config, str(ckpt))
    inf.model.set_compute_dtype("f32")
    prob, boxes = inf.infer_volume(val_cases[0][0], threshold=metrics["best_threshold"])
    assert prob.shape == val_cases[0][0].shape and len(boxes) >= 1
    dice = 2 * ((prob >= metrics["best_threshold"]) * val_cases[0][1]).sum() / ((prob >= metrics["best_threshold"]).sum() + val_cases[0][1].sum())
    print(f"inference on a validation case: {len(boxes)} boxes, Dice {dice:.3f}")
    assert dice > 0.3

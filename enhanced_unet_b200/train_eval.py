"""Drop-in for the hot-path callers of the reference ``train_eval.py``: ``Trainer`` (66-353), ``Evaluator``
(356-652, 852-1021 semantic part) and the ``train_model`` / ``evaluate_model`` drivers (1024-1232), with the
same names, constructor arguments, batch format and hyper-parameters.  Everything between the host loop and
the numbers is the C-ABI kernels: model forward/backward, fused loss, clip+AdamW, softmax/resize, the
probability->mask cascade and the confusion counts.

Out of scope: instance splitting (skimage), COCO evaluation, plotting and the
real-data ``CellDataset``; the drivers therefore accept any iterable of reference-format batches and ship a synthetic
generator.  The reference's per-image cv2 pre-processing in front of ``predict_semantic_mask`` (CLAHE + sharpening,
train_eval.py:365-395) is host-side image I/O, not arithmetic of the path: ``Evaluator._prepare_image_tensor`` makes the
same cv2 calls so that the entry point behaves like the reference's; everything after it runs on the GPU kernels.
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import metrics as _metrics
from . import parallel as _parallel
from .lib import call, ptr
from .models import DEFAULT_DTYPE, get_model
from .ops import combined_loss
from .optim import ClippedAdamW


class FocalLoss(torch.nn.Module):
    """Reference train_eval.py:28-60 with the enhanced_unet constants (alpha=[1,8,5], gamma=5,
    class_weights=[1,20,10]) - the focal term only, evaluated by torch ops for API compatibility.  The
    training path uses the fused kernel (``ops.combined_loss``), which contains this term."""

    def __init__(self, alpha=None, gamma=2.0, ignore_index=None, class_weights=None):
        super().__init__()
        self.alpha, self.gamma, self.ignore_index, self.class_weights = alpha, gamma, ignore_index, class_weights

    def forward(self, inputs, targets):
        kw = {} if self.ignore_index is None else {"ignore_index": self.ignore_index}
        ce = F.cross_entropy(inputs, targets, reduction="none", weight=self.class_weights, **kw)
        pt = torch.exp(-ce)
        if self.alpha is None:
            return ((1 - pt) ** self.gamma * ce).mean()
        if isinstance(self.alpha, (list, tuple, torch.Tensor)):
            a = torch.as_tensor(self.alpha, dtype=ce.dtype, device=ce.device)[targets]
        else:
            a = self.alpha
        return (a * (1 - pt) ** self.gamma * ce).mean()


def _pad32(images: torch.Tensor):
    """Reflect-pad [B,C,h,w] at the bottom / right to multiples of 32 (train_eval.py:249-253 / 400-406) on the pad kernel."""
    h, w = images.shape[-2:]
    h_pad, w_pad = (32 - h % 32) % 32, (32 - w % 32) % 32
    if h_pad or w_pad:
        if not images.is_cuda:
            raise RuntimeError("_pad32 runs on CUDA tensors (no CPU fallback)")
        if h_pad >= h or w_pad >= w:
            raise RuntimeError(f"reflect padding ({h_pad}, {w_pad}) must be smaller than the image ({h}, {w})")   # as F.pad
        src = images.contiguous().float()
        b, c = src.shape[:2]
        out = torch.empty(b, c, h + h_pad, w + w_pad, device=src.device, dtype=torch.float32)
        done, planes = 0, b * c
        sv, ov = src.view(planes, h, w), out.view(planes, h + h_pad, w + w_pad)
        while done < planes:
            k = min(65535, planes - done)
            call("eunet_reflect_pad", ptr(sv[done:]), ptr(ov[done:]), k, h, w, h + h_pad, w + w_pad)
            done += k
        images = out
    return images, h_pad, w_pad


class Trainer:
    """Reference train_eval.py:63-353 for model_name == 'enhanced_unet'."""

    def __init__(self, model, device, model_name, total_epochs: int = 50):
        if model_name != "enhanced_unet":
            raise NotImplementedError("only 'enhanced_unet' is on the B200 hot path")
        self.model, self.device, self.model_name = model, device, model_name
        self.total_epochs = max(1, total_epochs)
        self.dice_loss_weight, self.focal_loss_weight, self.tversky_loss_weight = 2.5, 2.5, 1.0   # train_eval.py:83-85
        base_lr = 4e-3                                                                             # train_eval.py:112
        on_update = getattr(getattr(model, "_packs", None), "invalidate", None)
        self.optimizer = ClippedAdamW(model.parameters(), lr=base_lr, weight_decay=1e-4, betas=(0.9, 0.999), max_norm=1.0,
                                      on_update=on_update)                                         # 120 + 341
        self.warmup_epochs = max(1, min(5, self.total_epochs // 6))
        self.scheduler = torch.optim.lr_scheduler.CosineAnnealingWarmRestarts(
            self.optimizer, T_0=max(10, self.total_epochs // 3), T_mult=2, eta_min=1e-7)
        self.warmup_scheduler = torch.optim.lr_scheduler.LinearLR(
            self.optimizer, start_factor=0.001, end_factor=1.0, total_iters=self.warmup_epochs)
        # data parallel (one process per GPU, torch.distributed initialised by the launcher; the reference has no
        # distributed code): identical replicas, gradients summed across ranks during backward, averaged in the optimiser
        self._world = _parallel.world_size()
        self._exchange = None
        if self._world > 1:
            _parallel.broadcast_parameters(list(model.parameters()) + list(model.buffers()))
            self._exchange = _parallel.GradientAllReduce(model)

    def _compute_combined_loss(self, logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        """One sample: logits [3,H,W] (or [3,2H,2W]), target [H,W] -> 2.5*focal + 2.5*dice + 1.0*tversky (183-197)."""
        return combined_loss(logits.unsqueeze(0), target.to(self.device).long().unsqueeze(0))

    def train_step(self, images: torch.Tensor, semantic_masks: List[torch.Tensor]) -> torch.Tensor:
        """One optimisation step on a reference-format batch; returns the (device) loss tensor."""
        images = images.to(self.device, non_blocking=True)
        self.optimizer.zero_grad(set_to_none=True)
        batch_size = images.shape[0]
        images, h_pad, w_pad = _pad32(images)
        outputs = self.model(images)
        shapes = {tuple(m.shape[-2:]) for m in semantic_masks}
        if len(shapes) == 1:
            # fast path: every mask has the image size -> one fused loss launch over the batch
            t = torch.stack([m.reshape(m.shape[-2:]) for m in semantic_masks]).to(self.device, non_blocking=True).long()
            if h_pad or w_pad:
                t = F.pad(t, (0, w_pad, 0, h_pad), mode="constant", value=0)                       # train_eval.py:281-288
            loss = combined_loss(outputs, t)
        else:
            loss = 0.0
            for i in range(batch_size):                                                            # train_eval.py:262-335
                gt = semantic_masks[i].reshape(semantic_masks[i].shape[-2:]).to(self.device).long()
                if h_pad or w_pad:
                    gt = F.pad(gt, (0, w_pad, 0, h_pad), mode="constant", value=0)
                loss = loss + combined_loss(outputs[i:i + 1], gt.unsqueeze(0))
            loss = loss / batch_size
        loss.backward()                # world > 1: launches the bucketed gradient all-reduce as gradients are finished
        if self._exchange is not None:
            self._exchange.wait()
        self.optimizer.step(grad_scale=1.0 / self._world)   # global-norm clip (max_norm=1.0) fused into the AdamW kernel
        return loss.detach()

    def train_epoch(self, dataloader: Iterable[Dict]) -> float:
        """Reference train_eval.py:236-353.  Batches: {'images': [B,3,H,W], 'batch_items': [{'semantic_mask': [H,W]}, ...]}."""
        self.model.train()
        losses = []
        for batch in dataloader:
            masks = [item["semantic_mask"] for item in batch["batch_items"]]
            losses.append(self.train_step(batch["images"], masks))
        if not losses:
            return 0.0
        return float(torch.stack(losses).mean())       # ONE device->host sync per epoch (reference: one per batch)


class Evaluator:
    """Reference train_eval.py:356-1021, semantic-segmentation part."""

    def __init__(self, model, device, model_name, tta: Optional[bool] = None):
        self.model, self.device, self.model_name = model, device, model_name
        # reference train_eval.py:363: enhanced_unet always averages the five TTA views; ``tta=False`` opts out
        self.enable_tta = (model_name == "enhanced_unet") if tta is None else bool(tta)

    @property
    def tta(self) -> bool:
        return self.enable_tta

    @tta.setter
    def tta(self, value: bool) -> None:
        self.enable_tta = bool(value)

    def _prepare_image_tensor(self, image: torch.Tensor) -> torch.Tensor:
        """Reference train_eval.py:365-395 (host-side, cv2): [3,h,w] in [0,1] (or 0..255) -> uint8 RGB -> CLAHE(2.0, 8x8) on
        the L channel of LAB -> 3x3 sharpening (0.15 * [[-1,-1,-1],[-1,9,-1],[-1,-1,-1]]) -> float [0,1] on the device."""
        try:
            import cv2
        except ImportError as e:      # loud: the reference cannot run without cv2 either
            raise RuntimeError("Evaluator._prepare_image_tensor needs OpenCV (cv2), as the reference does") from e
        host = image.detach().float().cpu()
        rgb = host.permute(1, 2, 0).numpy()
        rgb = (rgb * 255).astype(np.uint8) if float(host.max()) <= 1.0 else rgb.astype(np.uint8)
        lab = cv2.cvtColor(rgb, cv2.COLOR_RGB2LAB)
        lum, a_ch, b_ch = cv2.split(lab)
        lum = cv2.createCLAHE(clipLimit=2.0, tileGridSize=(8, 8)).apply(lum)
        rgb = cv2.cvtColor(cv2.merge([lum, a_ch, b_ch]), cv2.COLOR_LAB2RGB)
        sharpen = np.array([[-1, -1, -1], [-1, 9, -1], [-1, -1, -1]]) * 0.15
        rgb = np.clip(cv2.filter2D(rgb, -1, sharpen), 0, 255).astype(np.uint8)
        return torch.from_numpy(rgb.astype(np.float32) / 255.0).permute(2, 0, 1).to(self.device)

    @torch.no_grad()
    def _run_model_batch(self, images: torch.Tensor) -> torch.Tensor:
        """[B,3,h,w] -> probabilities [B,3,h,w] (pad to /32, model, bilinear resize == 2x2 mean, softmax; 397-417)."""
        b, _, h, w = images.shape
        x, h_pad, w_pad = _pad32(images.to(self.device))
        hp, wp = h + h_pad, w + w_pad
        logits = self.model(x)
        probs = torch.empty(b, 3, hp, wp, device=logits.device, dtype=torch.float32)
        call("eunet_softmax_probs", ptr(logits), ptr(probs), b, hp, wp, 2)
        return probs[:, :, :h, :w]

    def _run_model_single(self, image: torch.Tensor) -> torch.Tensor:
        return self._run_model_batch(image.unsqueeze(0))[0]

    @staticmethod
    def _resize(x: torch.Tensor, size=None, scale_factor: float = None) -> torch.Tensor:
        """``F.interpolate(x, size= | scale_factor=, mode='bilinear', align_corners=False)`` for [B,C,h,w] fp32 on the
        resize kernel, with ATen's output-size (floor(in * scale)) and source-ratio conventions."""
        import math
        b, c, h, w = x.shape
        if scale_factor is not None:
            ho, wo = int(math.floor(h * scale_factor)), int(math.floor(w * scale_factor))
            ratio = float(np.float32(1.0 / scale_factor))
            rh = rw = ratio
        else:
            ho, wo = size
            rh, rw = float(np.float32(h) / np.float32(ho)), float(np.float32(w) / np.float32(wo))
        x = x.contiguous().float()
        out = torch.empty(b, c, ho, wo, device=x.device, dtype=torch.float32)
        done = 0
        xs, os_ = x.view(b * c, h, w), out.view(b * c, ho, wo)
        while done < b * c:
            k = min(65535, b * c - done)
            call("eunet_resize_bilinear", ptr(xs[done:]), ptr(os_[done:]), k, h, w, ho, wo, rh, rw)
            done += k
        return out

    @torch.no_grad()
    def _run_tta_batch(self, images: torch.Tensor) -> torch.Tensor:
        """Reference _run_tta_inference (train_eval.py:419-453) for a batch: mean of the base view, the horizontal and
        vertical flips (index remaps) and the 0.75x / 1.25x bilinear views resized back to the input size."""
        images = images.to(self.device)
        b, _, h, w = images.shape
        views = [self._run_model_batch(images).contiguous(),
                 self._run_model_batch(images.flip(-1)).contiguous(),        # un-flipped by index inside the combine kernel
                 self._run_model_batch(images.flip(-2)).contiguous()]
        for scale in (0.75, 1.25):
            scaled = self._resize(images, scale_factor=scale)
            views.append(self._resize(self._run_model_batch(scaled), size=(h, w)))
        out = torch.empty(b, 3, h, w, device=views[0].device, dtype=torch.float32)
        done, planes = 0, b * 3
        vs = [v.view(planes, h, w) for v in views]
        ov = out.view(planes, h, w)
        while done < planes:
            k = min(65535, planes - done)
            call("eunet_tta_combine", *[ptr(v[done:]) for v in vs], ptr(ov[done:]), k, h, w)
            done += k
        return out

    def _run_tta_inference(self, image: torch.Tensor) -> torch.Tensor:
        """Single-image form with the reference's signature ([3,h,w] -> [3,h,w] probabilities)."""
        if not self.enable_tta:
            return self._run_model_single(image)
        return self._run_tta_batch(image.unsqueeze(0))[0]

    @torch.no_grad()
    def _probs(self, images: torch.Tensor) -> torch.Tensor:
        p = self._run_tta_batch(images) if self.enable_tta else self._run_model_batch(images)
        return p.contiguous()

    @torch.no_grad()
    def _convert_probs_to_mask_device(self, probs: torch.Tensor) -> torch.Tensor:
        """[B,3,h,w] probabilities -> uint8 masks [B,h,w] on the device (train_eval.py:455-568)."""
        probs = probs.contiguous().float()
        b, _, h, w = probs.shape
        mask = torch.empty(b, h, w, device=probs.device, dtype=torch.uint8)
        counts = torch.empty(b, 2, device=probs.device, dtype=torch.int32)
        call("eunet_probs_to_mask", ptr(probs), ptr(mask), ptr(counts), b, h, w)
        return mask

    def _convert_probs_to_mask(self, probs: torch.Tensor, h_pad: int = 0, w_pad: int = 0, h_orig: int = None,
                               w_orig: int = None) -> np.ndarray:
        m = self._convert_probs_to_mask_device(probs.unsqueeze(0).to(self.device))[0]
        if h_pad > 0 or w_pad > 0:
            m = m[:h_orig, :w_orig]
        return m.cpu().numpy().astype(np.int64)

    def predict_semantic_mask(self, image: torch.Tensor, preprocess: bool = True) -> np.ndarray:
        """[3,H,W] -> int64 [H,W] (reference 570-652: cv2 pre-processing -> TTA probabilities -> mask cascade).  CUDA errors
        propagate: the reference's silent retry on the CPU (576-592) is deliberately not reproduced.  ``preprocess=False``
        skips the host-side cv2 step (tensors that were pre-processed by the loader already)."""
        self.model.eval()
        if preprocess:
            image = self._prepare_image_tensor(image)
        probs = self._probs(image.unsqueeze(0))
        return self._convert_probs_to_mask_device(probs)[0].cpu().numpy().astype(np.int64)

    @torch.no_grad()
    def evaluate(self, dataloader: Iterable[Dict]) -> Dict[str, float]:
        """Mean of the per-image semantic metrics (reference 852-1021: per-image calculate_semantic_metrics at
        905, np.mean per key at 1014-1019).  The whole batch is predicted and counted on the device; one
        device->host copy of 16 counters per image."""
        self.model.eval()
        per_image: List[Dict] = []
        for batch in dataloader:
            images = batch["images"]
            gts = [item["semantic_mask"] for item in batch["batch_items"]]
            masks = self._convert_probs_to_mask_device(self._probs(images))
            gt = torch.stack([g.reshape(g.shape[-2:]) for g in gts]).to(masks.device).to(torch.uint8)
            per_image += _metrics.batch_semantic_metrics(masks, gt)
        if not per_image:
            return {}
        return {k: float(np.mean([float(m[k]) for m in per_image])) for k in per_image[0]}


# ---------------------------------------------------------------------------------------------
# drivers (reference train_eval.py:1024-1232) on any iterable of batches; synthetic data for the benchmarks
# ---------------------------------------------------------------------------------------------
class SyntheticCellBatches:
    """Iterable of reference-format batches of synthetic bright-field images."""

    def __init__(self, n_batches: int, batch_size: int, size: int, seed: int = 1234, device: str = "cpu"):
        self.n, self.b, self.s, self.seed, self.device = n_batches, batch_size, size, seed, device

    def __len__(self):
        return self.n

    def __iter__(self):
        for i in range(self.n):
            g = torch.Generator(device=self.device).manual_seed(self.seed + i)
            s = self.s
            yy = torch.arange(s, device=self.device, dtype=torch.float32).view(1, 1, s, 1)
            xx = torch.arange(s, device=self.device, dtype=torch.float32).view(1, 1, 1, s)
            nb = max(1, s * s // 4096)
            img = 0.75 + 0.05 * torch.randn(self.b, 1, s, s, device=self.device, generator=g)
            cy, cx = (torch.rand(self.b, nb, device=self.device, generator=g) * s for _ in range(2))
            sig = 3 + 7 * torch.rand(self.b, nb, device=self.device, generator=g)
            amp = (0.15 + 0.30 * torch.rand(self.b, nb, device=self.device, generator=g)).view(self.b, nb, 1, 1)
            d2 = (yy - cy.view(self.b, nb, 1, 1)) ** 2 + (xx - cx.view(self.b, nb, 1, 1)) ** 2
            blob = amp * torch.exp(-d2 / (2 * sig.view(self.b, nb, 1, 1) ** 2))
            img = (img - blob.sum(1, keepdim=True)).clamp_(0, 1).expand(self.b, 3, s, s).contiguous()
            label = (torch.where(amp < 0.3, 1, 2) * (blob > 0.5 * amp)).amax(1)
            yield {"images": img, "batch_items": [{"semantic_mask": label[j]} for j in range(self.b)]}


def train_model(model_name: str, data_dir, device: str = "cuda", num_epochs: int = 50, skip_training: bool = False,
                train_batches: Optional[Iterable[Dict]] = None, val_batches: Optional[Iterable[Dict]] = None,
                dtype: str = DEFAULT_DTYPE, tta: Optional[bool] = None) -> str:
    """Reference train_eval.py:1036-1162: epochs with warm-up / cosine-restart LR, validation every 3rd epoch (with the
    5-view TTA, as the reference's Evaluator does for enhanced_unet; ``tta=False`` opts out), best-mIoU checkpoint in the
    reference's format, early stopping (patience 10 validations without a better mIoU, only after epoch 26).
    ``data_dir`` is unused for synthetic runs."""
    save_dir = os.path.join("checkpoints", model_name)
    os.makedirs(save_dir, exist_ok=True)
    checkpoint_path = os.path.join(save_dir, "best_model.pth")
    if os.path.exists(checkpoint_path) and skip_training:
        return checkpoint_path
    train_batches = train_batches if train_batches is not None else SyntheticCellBatches(8, 2, 256, seed=1)
    val_batches = val_batches if val_batches is not None else SyntheticCellBatches(2, 2, 256, seed=99)
    model = get_model(model_name, num_classes=3, device=device, dtype=dtype).to(device)
    trainer = Trainer(model, device, model_name, total_epochs=num_epochs)
    history = {"train_loss": [], "val_miou": [], "learning_rate": [], "epoch_axis": []}
    best_miou, best_loss = 0.0, float("inf")                                                         # 1095-1096
    patience, patience_counter = 10, 0                                                               # 1097-1098
    for epoch in range(num_epochs):
        (trainer.warmup_scheduler if epoch < trainer.warmup_epochs else trainer.scheduler).step()      # 1104-1111
        lr = trainer.optimizer.param_groups[0]["lr"]
        loss = trainer.train_epoch(train_batches)
        history["train_loss"].append(loss)
        history["learning_rate"].append(lr)
        print(f"Epoch {epoch + 1}/{num_epochs}  lr {lr:.6f}  loss {loss:.4f}")
        if (epoch + 1) % 3 == 0:                                                                     # 1118
            res = Evaluator(model, device, model_name, tta=tta).evaluate(val_batches)
            miou = res.get("sem_mean_iou", 0.0)
            history["val_miou"].append(miou)
            history["epoch_axis"].append(epoch + 1)
            print(f"  val mIoU {miou:.4f}  live {res.get('sem_live_iou', 0):.4f}  dead {res.get('sem_dead_iou', 0):.4f}")
            if miou > best_miou:
                best_miou, best_loss, patience_counter = miou, loss, 0
                if _parallel.world_size() == 1 or torch.distributed.get_rank() == 0:                 # rank 0's replica is checkpointed
                    torch.save({"epoch": epoch + 1, "model_state_dict": model.state_dict(),
                                "optimizer_state_dict": trainer.optimizer.state_dict(),
                                "scheduler_state_dict": trainer.scheduler.state_dict(), "best_miou": best_miou,
                                "best_loss": best_loss, "history": history}, checkpoint_path)        # 1143-1151
            else:
                patience_counter += 1
        if patience_counter >= patience and epoch > 25:                                              # 1156-1159
            print(f"Early stopping at epoch {epoch + 1}")
            break
    if hasattr(model, "check_numerics"):
        model.check_numerics()
    if not os.path.exists(checkpoint_path) and (_parallel.world_size() == 1 or torch.distributed.get_rank() == 0):
        # fewer than three epochs (or no validation ever beat mIoU 0): the reference would return a path that does not
        # exist; save the final weights so that evaluate_model has something to load
        torch.save({"epoch": num_epochs, "model_state_dict": model.state_dict(), "optimizer_state_dict": trainer.optimizer.state_dict(),
                    "scheduler_state_dict": trainer.scheduler.state_dict(), "best_miou": best_miou, "best_loss": best_loss,
                    "history": history}, checkpoint_path)
    return checkpoint_path


def evaluate_model(model_name: str, data_dir, device: str = "cuda", checkpoint_path: Optional[str] = None,
                   batches: Optional[Iterable[Dict]] = None, dtype: str = DEFAULT_DTYPE, tta: Optional[bool] = None) -> Dict[str, float]:
    """Reference train_eval.py:1165-1232 (semantic metrics; 5-view TTA for enhanced_unet unless ``tta=False``)."""
    model = get_model(model_name, num_classes=3, device=device, dtype=dtype)
    if checkpoint_path and os.path.exists(checkpoint_path):
        ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)                   # 1190-1191
        model.load_state_dict(ckpt["model_state_dict"])
    model = model.to(device)
    batches = batches if batches is not None else SyntheticCellBatches(2, 2, 256, seed=99)
    return Evaluator(model, device, model_name, tta=tta).evaluate(batches)


def train_and_evaluate(model_name: str, data_dir, device: str = "cuda", num_epochs: int = 50, **kw) -> Dict[str, float]:
    """Reference train_eval.py:1024-1033."""
    ckpt = train_model(model_name, data_dir, device, num_epochs, **kw)
    return evaluate_model(model_name, data_dir, device, ckpt, dtype=kw.get("dtype", DEFAULT_DTYPE), tta=kw.get("tta"))

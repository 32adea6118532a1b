"""CLI mirroring the reference ``main.py`` (flags of main.py:86-96) for the accelerated path.

    python -m enhanced_unet_b200.main --mode train_eval --models enhanced_unet --epochs 3 --synthetic

``--synthetic`` (default; the reference's labelme ``data/`` directory and cv2 pre-processing are out of scope)
generates bright-field-like batches on the fly.  Extra flags: ``--dtype {fp16,bf16,fp32}``, ``--size``, ``--batch``."""
from __future__ import annotations

import argparse
import json
import os

import torch

from .train_eval import SyntheticCellBatches, evaluate_model, train_model


def main(argv=None):
    ap = argparse.ArgumentParser(description="Enhanced-UNet (B200 hot path)")
    ap.add_argument("--mode", default="train_eval", choices=["train", "eval", "train_eval", "visualize"])
    ap.add_argument("--models", nargs="+", default=["enhanced_unet"])
    ap.add_argument("--epochs", type=int, default=50)
    ap.add_argument("--regenerate-predictions", action="store_true", help="accepted for compatibility; unused")
    ap.add_argument("--synthetic", action="store_true", default=True)
    ap.add_argument("--dtype", default="fp16", choices=["fp16", "bf16", "fp32"])
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--train-batches", type=int, default=8)
    args = ap.parse_args(argv)
    if args.mode == "visualize":
        raise SystemExit("plotting (reference visualization.py) is outside the accelerated hot path")
    if not torch.cuda.is_available():
        raise SystemExit("a CUDA device is required (there is no CPU fallback)")
    # one process per GPU under torchrun (WORLD_SIZE > 1): NCCL process group, rank-local batches, gradient all-reduce
    # inside Trainer; rank 0 evaluates and writes the results
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    if world > 1 and not torch.distributed.is_initialized():
        torch.distributed.init_process_group("nccl", device_id=torch.device("cuda", local))
    device = "cuda" if world == 1 else f"cuda:{local}"
    all_results = {}
    for name in args.models:
        train = SyntheticCellBatches(args.train_batches, args.batch, args.size, seed=1 + 1000 * rank)
        val = SyntheticCellBatches(2, args.batch, args.size, seed=99)
        ckpt = os.path.join("checkpoints", name, "best_model.pth")
        if args.mode in ("train", "train_eval"):
            ckpt = train_model(name, "data", device, args.epochs, train_batches=train, val_batches=val, dtype=args.dtype)
        if args.mode in ("eval", "train_eval") and rank == 0:
            all_results[name] = evaluate_model(name, "data", device, ckpt, batches=val, dtype=args.dtype)
            print(json.dumps({name: all_results[name]}, indent=1))
    if rank == 0:
        os.makedirs("results", exist_ok=True)
        with open(os.path.join("results", "evaluation_results.json"), "w") as f:      # reference main.py:251-279
            json.dump(all_results, f, indent=2)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return all_results


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Where does the whole-model step spend its time?  (GPU box; evidence for bench.py's `model_e2e` numbers.)

Runs the DETRPose-L training step (or DETRPose-S/X inference) of baseline/model_bench.py under torch.profiler
for a few steps, with and without the kernels installed, and prints one JSON line per arm:
device-busy time per step (sum of kernel durations), wall time per step, kernels per step, and the device time
per group of kernels (this package's kernels, cuDNN/cuBLAS convolutions and GEMMs, the reference's grid_sample
path, NCCL, everything else).  `device_busy_frac` well below 1 means the step is bound by the host (Python,
kernel launches, the Hungarian matcher on the CPU), not by any kernel.

    python tools/profile_model.py [--leg train_l|infer_s|infer_x] [--steps 3]
    torchrun --nproc-per-node 8 tools/profile_model.py --leg train_l      # rank 0 reports; DDP all-reduce included
"""
import argparse
import collections
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from baseline import ref_harness as rh       # noqa: E402

GROUPS = [
    ("msda_b200", ("msda::", "fwd_lean", "bwd_gather", "gate_fwd", "gate_bwd", "lqe_", "softmax_bwd", "repack",
                   "unpack_grad", "locations_kernel")),
    ("nccl", ("nccl",)),
    ("grid_sample(reference sampler)", ("grid_sampler",)),
    ("conv/gemm (cuDNN, cuBLAS, cutlass)", ("cudnn", "cutlass", "gemm", "sm90", "sm100", "xmma", "conv", "cublas",
                                           "nvjet", "wgrad", "dgrad", "fprop")),
    ("norm/softmax/attention", ("batch_norm", "layer_norm", "softmax", "fmha", "flash", "sdpa", "bn_fw", "bn_bw")),
    ("optimizer/foreach", ("multi_tensor", "adam", "foreach")),
]


def group_of(name: str) -> str:
    low = name.lower()
    for g, keys in GROUPS:
        if any(k in low for k in keys):
            return g
    return "elementwise/copy/other"


def build_step(leg, dev, world, sync_bn=True, static_graph=False):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    if leg == "train_l":
        model = rh.build_model("l", seed=0).to(dev).train()
        criterion = rh.build_criterion().to(dev).train()
        net = model
        if world > 1:
            if sync_bn:                                   # src/misc/dist_utils.py:122 (sync_bn: True in the configs)
                model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
            if static_graph:                              # what-if: not what the reference does (:126)
                net = DDP(model, device_ids=[dev.index], output_device=dev.index, static_graph=True,
                          gradient_as_bucket_view=True)
            else:
                net = DDP(model, device_ids=[dev.index], output_device=dev.index, find_unused_parameters=True)
        params = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.AdamW(params, lr=1e-4, weight_decay=1e-4)
        host = torch.rand(16, 3, 640, 640).pin_memory()
        targets = rh.synthetic_targets(16, dev, seed=dist.get_rank() if world > 1 else 0)

        def step():
            images = host.to(dev, non_blocking=True)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = net(images, targets)
            with torch.autocast("cuda", enabled=False):
                loss = sum(criterion(out, targets).values()) + model.layer_loss.to(dev)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 0.1)
            opt.step()
            opt.zero_grad()
            return loss
        return step, 16
    size, batch = ("s", 64) if leg == "infer_s" else ("x", 32)
    model = rh.build_model(size, seed=0).to(dev).deploy()
    host = torch.rand(batch, 3, 640, 640).pin_memory()

    def step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
            return model(host.to(dev, non_blocking=True))["pred_logits"]
    return step, batch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--leg", default="train_l", choices=["train_l", "infer_s", "infer_x"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--arms", default="reference,b200")
    ap.add_argument("--no-syncbn", action="store_true", help="what-if: plain BatchNorm under DDP (the reference converts to SyncBN)")
    ap.add_argument("--static-graph", action="store_true", help="what-if: DDP(static_graph=True, gradient_as_bucket_view=True)")
    ap.add_argument("--no-profiler", action="store_true", help="time only (CUDA events), no torch.profiler")
    args = ap.parse_args()
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    devnull = open(os.devnull, "w")
    for arm in args.arms.split(","):
        rh.uninstall_kernels()
        if arm == "b200":
            rh.install_kernels()
        old, sys.stdout = sys.stdout, devnull
        try:
            step, batch = build_step(args.leg, dev, world, sync_bn=not args.no_syncbn, static_graph=args.static_graph)
        finally:
            sys.stdout = old
        for _ in range(3):
            step()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        if args.no_profiler:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / args.steps
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            if rank == 0:
                print(json.dumps({"leg": args.leg, "arm": arm, "n_gpus": world, "batch_per_gpu": batch,
                                  "sync_bn": not args.no_syncbn, "static_graph": args.static_graph,
                                  "ms_per_step": round(ms, 2), "img_per_s": round(batch * world / ms * 1e3, 1)}), flush=True)
            del step
            torch.cuda.empty_cache()
            continue
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step()
            torch.cuda.synchronize(dev)
            wall = (time.perf_counter() - t0) / args.steps
        per_group = collections.Counter()
        per_kernel = collections.Counter()
        launches = 0
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                d = ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
                per_group[group_of(ev.name)] += d
                per_kernel[ev.name[:70]] += d
                launches += 1
        busy = sum(per_group.values()) / args.steps / 1e3
        if rank == 0:
            print(json.dumps({
                "leg": args.leg, "arm": arm, "n_gpus": world, "batch_per_gpu": batch,
                "wall_ms_per_step": round(wall * 1e3, 2), "device_busy_ms_per_step": round(busy, 2),
                "device_busy_frac": round(busy / (wall * 1e3), 3), "kernels_per_step": launches // args.steps,
                "device_ms_per_step_by_group": {k: round(v / args.steps / 1e3, 2) for k, v in per_group.most_common()},
                "top_kernels_ms_per_step": {k: round(v / args.steps / 1e3, 2) for k, v in per_kernel.most_common(8)},
            }), flush=True)
        del step
        torch.cuda.empty_cache()
    rh.uninstall_kernels()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

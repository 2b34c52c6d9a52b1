"""Whole-model legs of ``bench.py`` (key ``model_e2e``): the UNMODIFIED reference DETRPose with this
package's kernels dropped in, beside the same model on the reference's own PyTorch path, on the same GPU.

BASELINE.json configs[1..3]:
  * ``infer_s``  DETRPose-S inference 640x640, batch 64, bf16 autocast, ``deploy()+eval()`` exactly as
    tools/benchmark/torch_benchmark.py:82-93 builds it (post-processor included);
  * ``train_l``  DETRPose-L training step, batch 16/GPU, OKS-denoising queries, the step of
    src/solver/engine.py:37-89 (autocast forward, criterion in fp32, backward, clip 0.1, AdamW), under
    ``DistributedDataParallel(find_unused_parameters=True)`` + SyncBN as src/misc/dist_utils.py:111-133
    wraps it when more than one rank runs -- the gradient all-reduce over NCCL is inside the timed step;
  * ``infer_x``  DETRPose-X inference, batch 32/GPU (256 over 8 GPUs), bf16 autocast.
Every step uploads its images from pinned host memory and reads a result back (top score / loss), both
inside the timed region.  Each leg is timed twice: ``reference`` (nothing installed) and ``b200``
(``ref_harness.install_kernels()``: core + value hand-over + gate + LQE).  Random-init weights, synthetic
images and targets; img/s = images of all ranks / max-over-ranks device time.
"""
from __future__ import annotations

import contextlib
import time

import torch

from . import ref_harness as rh


def _max_over_ranks(x, dev):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return float(x)


def _barrier(dev):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    torch.cuda.synchronize(dev)


def _timed(step, steps, warmup, dev):
    """``steps`` calls of ``step()`` between CUDA events (barrier + synchronize on both sides), max over ranks."""
    for _ in range(warmup):
        step()
    _barrier(dev)
    stream = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(steps):
        step()
    e1.record(stream)
    _barrier(dev)
    wall = (time.perf_counter() - t0) * 1e3
    return _max_over_ranks(e0.elapsed_time(e1), dev) / steps, wall / steps


def _inference_leg(size, batch, dev, world, steps, warmup, dtype=torch.bfloat16):
    net = rh.build_model(size, seed=0).to(dev)
    post = rh.build_postprocessor().to(dev)

    class Model(torch.nn.Module):                     # tools/benchmark/torch_benchmark.py:82-93
        def __init__(self):
            super().__init__()
            self.model = net.deploy()
            self.postprocessor = post.deploy()

        def forward(self, images, orig_target_sizes):
            outputs = self.model(images)
            if images.shape[0] == 1:
                return self.postprocessor(outputs, orig_target_sizes)[0]
            # PostProcess in deploy mode expands its gather index to batch 1 (postprocesses.py:27-30) and its
            # other branch expects the un-deployed (B, 60, 34) keypoint layout: for a batch, the same top-k
            # score selection is done here (postprocesses.py:19-21)
            prob = outputs["pred_logits"].sigmoid()
            return torch.topk(prob.view(prob.shape[0], -1), 60, dim=1)[0]

    model = Model().to(dev)
    host = torch.rand(batch, 3, 640, 640).pin_memory()
    sizes = torch.tensor([[640, 640]] * batch, device=dev)
    sink = torch.empty(batch, 60).pin_memory()

    def step():
        with torch.no_grad(), torch.autocast("cuda", dtype=dtype):
            images = host.to(dev, non_blocking=True)
            scores = model(images, sizes)
            sink.copy_(scores.float(), non_blocking=True)

    res = {}
    for arm in ("reference", "b200"):
        rh.uninstall_kernels()
        if arm == "b200":
            rh.install_kernels()
        ms, wall = _timed(step, steps, warmup, dev)
        res[arm] = {"ms_per_step": round(ms, 3), "img_per_s": round(batch * world / ms * 1e3, 1),
                    "host_wall_ms_per_step": round(wall, 3)}
    rh.uninstall_kernels()
    res["speedup"] = round(res["b200"]["img_per_s"] / res["reference"]["img_per_s"], 3)
    res["config"] = (f"DETRPose-{size.upper()} deploy()+eval()+top-k scores, 640x640, batch {batch}/GPU, "
                     f"{str(dtype).split('.')[-1]} autocast, random-init weights, images uploaded from pinned "
                     f"host memory and top scores read back every step")
    res["h2d_bytes_per_step"] = host.numel() * 4
    res["d2h_bytes_per_step"] = sink.numel() * 4
    del model, net
    torch.cuda.empty_cache()
    return res


def _training_leg(size, batch, dev, world, steps, warmup, dtype=torch.bfloat16):
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    res = {}
    host = torch.rand(batch, 3, 640, 640).pin_memory()
    targets = rh.synthetic_targets(batch, dev, seed=dist.get_rank() if dist.is_initialized() else 0)
    for arm in ("reference", "b200"):
        rh.uninstall_kernels()
        if arm == "b200":
            rh.install_kernels()
        model = rh.build_model(size, seed=0).to(dev).train()
        criterion = rh.build_criterion().to(dev).train()
        ddp = None
        if dist.is_initialized() and world > 1:       # src/misc/dist_utils.py:119-126
            model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
            ddp = DDP(model, device_ids=[dev.index], output_device=dev.index, find_unused_parameters=True)
        net = ddp if ddp is not None else model
        params = [p for p in model.parameters() if p.requires_grad]
        opt = torch.optim.AdamW(params, lr=1e-4, betas=(0.9, 0.999), weight_decay=1e-4)
        sink = torch.empty(1).pin_memory()

        def step():                                    # src/solver/engine.py:37-89, grad_accum_steps = 1
            images = host.to(dev, non_blocking=True)
            with torch.autocast("cuda", dtype=dtype):
                outputs = net(images, targets)
            with torch.autocast("cuda", enabled=False):
                loss_dict = criterion(outputs, targets)
                loss = sum(loss_dict.values()) + model.layer_loss.to(dev)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(params, 0.1)
            opt.step()
            opt.zero_grad()
            sink.copy_(loss.detach().reshape(1), non_blocking=True)

        ms, wall = _timed(step, steps, warmup, dev)
        torch.cuda.synchronize(dev)
        res[arm] = {"ms_per_step": round(ms, 3), "img_per_s": round(batch * world / ms * 1e3, 1),
                    "host_wall_ms_per_step": round(wall, 3), "loss": round(float(sink.item()), 4),
                    "peak_mem_GB": round(torch.cuda.max_memory_allocated(dev) / 2 ** 30, 2)}
        del net, ddp, model, criterion, opt
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats(dev)
    rh.uninstall_kernels()
    res["speedup"] = round(res["b200"]["img_per_s"] / res["reference"]["img_per_s"], 3)
    res["config"] = (f"DETRPose-{size.upper()} training step (engine.py:37-89: autocast-"
                     f"{str(dtype).split('.')[-1]} forward with OKS-denoising queries, fp32 criterion with "
                     f"Hungarian matching on the host, backward, clip 0.1, AdamW), batch {batch}/GPU, "
                     f"grad_accum_steps 1, synthetic targets (1-8 persons/image), "
                     + ("DDP(find_unused_parameters=True) + SyncBN over NCCL" if world > 1 else "single process"))
    res["h2d_bytes_per_step"] = host.numel() * 4
    res["d2h_bytes_per_step"] = 4
    return res


def run(dev, world: int, quick: bool = False) -> dict:
    """All legs; an exception in one leg is recorded in its entry and does not stop the others."""
    if not rh.available():
        return {"unavailable": "baseline/_ref (vendored reference sources) not present"}
    out = {}
    legs = [
        ("infer_s", lambda: _inference_leg("s", 64, dev, world, steps=3 if quick else 8, warmup=2)),
        ("train_l", lambda: _training_leg("l", 16, dev, world, steps=3 if quick else 6, warmup=2)),
        ("infer_x", lambda: _inference_leg("x", 32, dev, world, steps=3 if quick else 6, warmup=2)),
    ]
    for name, fn in legs:
        t0 = time.perf_counter()
        try:
            with contextlib.redirect_stdout(open("/dev/null", "w")):
                out[name] = fn()
        except Exception as e:                                    # noqa: BLE001 -- reported, not swallowed
            rh.uninstall_kernels()
            out[name] = {"error": f"{type(e).__name__}: {e}"[:400]}
            torch.cuda.empty_cache()
        out[name]["leg_wall_s"] = round(time.perf_counter() - t0, 1)
    return out

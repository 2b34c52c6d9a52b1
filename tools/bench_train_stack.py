"""Training step of the pose decoder's cross-attention blocks on N GPUs (replicas + DDP
gradient all-reduce over NCCL, as the reference trains -- src/misc/dist_utils.py:126).

The full DETRPose-L model cannot travel to the GPU box (the reference is not vendored), so this times the
part of its training step this repository covers, at the L training shape: 6 decoder layers, each

    query -> MSDeformAttn(query, reference_points, value, shapes)      (offset / attention Linears, prologue, core)
          -> Gate(query, sampled)                                      (2C->2C Linear, sigmoid/blend/LayerNorm)

on one shared `memory` (N, 8400, 256) that requires grad (the encoder's output in the real model), batch
16 per GPU, Lq = 1584 denoising + matching queries x 18 keypoints, loss = mean(out^2), backward, DDP
all-reduce of the parameter gradients, AdamW step.  `--impl reference` runs the same stack with the
reference's op sequence (per-level F.grid_sample + stack/mul/sum, cat + Linear + sigmoid/chunk/blend +
LayerNorm) restated in tools/sweep.py and tools/bench_gate.py.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port 29540 tools/bench_train_stack.py [--impl reference] [--autocast]

One JSON line from rank 0: images/s over all ranks (max over ranks of the device time of K steps).
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn.functional as F
from torch import nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import detrpose_b200 as dp                                             # noqa: E402
from detrpose_b200 import synthetic, shard                             # noqa: E402
from sweep import reference_ops                                        # noqa: E402


class RefLayer(nn.Module):
    """The reference's ops for one layer (ms_deform_attn.py:385-416, 145-193; transformer.py:231-235)."""

    def __init__(self, C, H, L, P):
        super().__init__()
        self.H, self.L, self.P = H, L, P
        self.sampling_offsets = nn.Linear(C, H * L * P * 2)
        self.attention_weights = nn.Linear(C, H * L * P)
        self.gate = nn.Linear(2 * C, 2 * C)
        self.norm = nn.LayerNorm(C)

    def forward(self, query, ref, memory, shapes):
        n, lq, c = query.shape
        H, L, P = self.H, self.L, self.P
        off = self.sampling_offsets(query).view(n, lq, H, L, P, 2)
        att = F.softmax(self.attention_weights(query).view(n, lq, H, L * P), -1).view(n, lq, H, L, P)
        norm = torch.tensor([[w, h] for h, w in shapes], device=query.device)
        loc = ref[:, :, None, :, None, :] + off / norm[None, None, None, :, None, :]
        S = memory.shape[1]
        v = memory.view(n, S, H, c // H).permute(0, 2, 3, 1).flatten(0, 1)
        vlist = list(v.split([h * w for h, w in shapes], dim=-1))
        sampled = reference_ops([t.float() for t in vlist], shapes, loc.float(), att.float()).to(query.dtype)
        gates = torch.sigmoid(self.gate(torch.cat([query, sampled], dim=-1)))
        g1, g2 = gates.chunk(2, dim=-1)
        return self.norm(g1 * query + g2 * sampled)


class OurLayer(nn.Module):
    def __init__(self, C, H, L, P):
        super().__init__()
        self.attn = dp.MSDeformAttn(d_model=C, n_levels=L, n_heads=H, n_points=P)
        self.gate = dp.Gate(C)

    def forward(self, query, ref, memory, shapes):
        # reference_points arrive as (N, nq, 1, K, 2) in the model; here nq = Lq, K = 1
        sampled = self.attn(query, ref.unsqueeze(3), memory, shapes)
        return self.gate(query, sampled.to(query.dtype))


class Stack(nn.Module):
    def __init__(self, impl, layers, C, H, L, P):
        super().__init__()
        cls = OurLayer if impl == "b200" else RefLayer
        self.layers = nn.ModuleList(cls(C, H, L, P) for _ in range(layers))

    def forward(self, query, ref, memory, shapes):
        for layer in self.layers:
            query = layer(query, ref, memory, shapes)
        return query


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--lq", type=int, default=1584)
    ap.add_argument("--layers", type=int, default=6)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--autocast", action="store_true", help="bf16 autocast (bf16 memory) instead of fp32")
    ap.add_argument("--no-fused-prologue", action="store_true", help="softmax / locations through torch ops")
    ap.add_argument("--graph", action="store_true",
                    help="capture forward + backward + AdamW of one step in a CUDA graph and replay it (1 GPU)")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    w = synthetic.WORKLOADS["detrpose_l"]
    C, H, P, shapes = w["H"] * w["Dh"], w["H"], w["P"], [list(s) for s in w["shapes"]]
    L, S = len(shapes), synthetic.pyramid_size(w["shapes"])
    torch.manual_seed(1234)                                   # identical initial replicas
    model = Stack(args.impl, args.layers, C, H, L, P).to(dev)
    with torch.no_grad():                                     # leave the zero init: spread the samples
        for m in model.modules():
            if isinstance(m, nn.Linear) and m.weight.abs().max() == 0:
                m.weight.normal_(0, 0.02)
    if args.no_fused_prologue:
        for m in model.modules():
            if isinstance(m, dp.MSDeformAttn):
                m.fuse_prologue = False
    ddp = nn.parallel.DistributedDataParallel(model, device_ids=[local]) if world > 1 else model
    opt = torch.optim.AdamW(ddp.parameters(), lr=1e-4, capturable=args.graph)

    g = torch.Generator(device=dev).manual_seed(100 + rank)   # each rank its own images
    mdt = torch.bfloat16 if args.autocast else torch.float32
    memory = torch.randn(args.batch, S, C, device=dev, generator=g).to(mdt).requires_grad_(True)
    query = torch.randn(args.batch, args.lq, C, device=dev, generator=g)
    ref = torch.rand(args.batch, args.lq, 1, 2, device=dev, generator=g)

    def step():
        opt.zero_grad(set_to_none=True)
        memory.grad = None
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=args.autocast):
            out = ddp(query, ref, memory, shapes)
            loss = out.float().square().mean()
        loss.backward()                                        # DDP all-reduces the parameter gradients here
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if args.graph:
        if world > 1:
            raise SystemExit("--graph is a single-GPU option")
        # The C-ABI calls only enqueue work on the current stream (no allocation, no synchronisation), so a
        # whole training step -- autograd included -- can be captured once and replayed.
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                step()
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        opt.zero_grad(set_to_none=True)
        memory.grad = None
        with torch.cuda.graph(graph):
            captured_loss = step()
        eager_step = step

        def step():                                            # noqa: F811 - the timed loop replays the graph
            graph.replay()
            return captured_loss

    for _ in range(args.warmup):
        step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    barrier()
    ms = shard.max_over_ranks(e0.elapsed_time(e1), device=dev)
    loss_v = float(loss.detach())
    if rank == 0:
        line = {"metric": "pose_decoder_cross_attention_blocks_train_images_per_s", "impl": args.impl,
                "value": round(world * args.batch * args.steps / (ms * 1e-3), 1), "unit": "img/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
                "scaling": "weak", "dtype": "bf16 autocast" if args.autocast else "fp32",
                "config": {"layers": args.layers, "images_per_gpu": args.batch, "Lq": args.lq, "d_model": C,
                           "levels": shapes, "points": P, "optimizer": "AdamW",
                           "gradient_sync": "DDP all-reduce (NCCL)" if world > 1 else "none (1 GPU)",
                           "cuda_graph": bool(args.graph)},
                "final_loss": round(loss_v, 6),
                "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}
        sys.stdout.write(json.dumps(line) + "\n")
        sys.stdout.flush()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""Generate the golden fixtures by running the REAL reference (build container only).

    python tests/golden/make_golden.py        # needs /root/reference, writes tests/golden/*.npz

The reference ships no tests or golden vectors of its own, so the
vectors committed next to this script are outputs of the reference's own
``ms_deform_attn_core_pytorch`` / ``MSDeformAttn`` (imported by file path from
/root/reference/src/models/detrpose/ms_deform_attn.py) on seeded inputs, in fp32
and -- as the high-precision arbiter -- in fp64 (stored rounded to fp32).
/root/reference does not exist on the GPU box; only these fixtures travel.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_FILE = "/root/reference/src/models/detrpose/ms_deform_attn.py"

CORE_CASES = {
    # name: N, H, Dh, Lq, P, shapes, location range, flavour
    "tiny_2lvl":     dict(N=2, H=2, Dh=8,  Lq=7,  P=3, shapes=((5, 7), (3, 4)), lo=-0.3, hi=1.3),
    "n_like":        dict(N=1, H=8, Dh=16, Lq=20, P=6, shapes=((10, 10), (5, 5)), lo=-0.1, hi=1.1),
    "s_like":        dict(N=2, H=8, Dh=32, Lq=18, P=4, shapes=((8, 8), (4, 4), (2, 2)), lo=-0.1, hi=1.1),
    "x_like":        dict(N=1, H=8, Dh=48, Lq=9,  P=4, shapes=((8, 10), (4, 5), (2, 3)), lo=-0.1, hi=1.1),
    "sweep4_like":   dict(N=1, H=8, Dh=32, Lq=10, P=4, shapes=((8, 8), (4, 4), (2, 2), (1, 1)), lo=0.0, hi=1.0),
    "pixel_centres": dict(N=1, H=2, Dh=8,  Lq=12, P=4, shapes=((4, 6), (2, 3)), flavour="centres"),
    "degenerate":    dict(N=2, H=4, Dh=16, Lq=6,  P=4, shapes=((6, 6), (3, 3)), lo=0.0, hi=1.0, flavour="degenerate"),
    "far_outside":   dict(N=1, H=2, Dh=8,  Lq=8,  P=2, shapes=((4, 4), (2, 2)), lo=-3.0, hi=4.0),
}


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_msda", REF_FILE)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def reference_value_list(memory, n_heads, shapes):
    """The caller's construction, transformer.py:1285-1286, verbatim semantics."""
    split_sizes = [h * w for h, w in shapes]
    value = memory.unflatten(2, (n_heads, -1))
    return value.permute(0, 2, 3, 1).flatten(0, 1).split(split_sizes, dim=-1)


def make_case_inputs(name, cfg, seed):
    g = torch.Generator().manual_seed(seed)
    N, H, Dh, Lq, P, shapes = cfg["N"], cfg["H"], cfg["Dh"], cfg["Lq"], cfg["P"], cfg["shapes"]
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    memory = torch.randn(N, S, H * Dh, generator=g)
    flavour = cfg.get("flavour")
    if flavour == "centres":
        # exactly on pixel centres, on the borders, and on cell corners
        loc = torch.empty(N, Lq, H, L, P, 2)
        for l, (h, w) in enumerate(shapes):
            ix = torch.randint(0, 2 * w + 1, (N, Lq, H, P), generator=g).float()
            iy = torch.randint(0, 2 * h + 1, (N, Lq, H, P), generator=g).float()
            loc[:, :, :, l, :, 0] = ix / (2 * w)       # multiples of half a pixel: centres and edges
            loc[:, :, :, l, :, 1] = iy / (2 * h)
        attn = torch.softmax(torch.randn(N, Lq, H, L * P, generator=g), -1).view(N, Lq, H, L, P)
    elif flavour == "degenerate":
        one = cfg["lo"] + (cfg["hi"] - cfg["lo"]) * torch.rand(N, Lq, H, L, 1, 2, generator=g)
        loc = one.expand(N, Lq, H, L, P, 2).contiguous()
        attn = torch.full((N, Lq, H, L, P), 1.0 / (L * P))
    else:
        loc = cfg["lo"] + (cfg["hi"] - cfg["lo"]) * torch.rand(N, Lq, H, L, P, 2, generator=g)
        attn = torch.softmax(torch.randn(N, Lq, H, L * P, generator=g), -1).view(N, Lq, H, L, P)
    grad_out = torch.randn(N, Lq, H * Dh, generator=g)
    return memory, loc, attn, grad_out


def run_reference_core(ref, memory, loc, attn, grad_out, n_heads, shapes, dtype):
    memory = memory.to(dtype).requires_grad_(True)
    loc = loc.to(dtype).requires_grad_(True)
    attn = attn.to(dtype).requires_grad_(True)
    value = reference_value_list(memory, n_heads, shapes)
    out = ref.ms_deform_attn_core_pytorch(value, [list(s) for s in shapes], loc, attn)
    g_mem, g_loc, g_attn = torch.autograd.grad(out, [memory, loc, attn], grad_out.to(dtype))
    return out.detach(), g_mem, g_loc, g_attn


def reference_indices(loc, shapes):
    """(y0, x0) via the reference's op chain in fp32: 2*loc-1 (:161), ((g+1)*size-1)/2, floor."""
    grid = 2 * loc - 1
    idx = torch.empty(loc.shape, dtype=torch.int32)
    for l, (h, w) in enumerate(shapes):
        x = ((grid[:, :, :, l, :, 0] + 1) * w - 1) / 2
        y = ((grid[:, :, :, l, :, 1] + 1) * h - 1) / 2
        idx[:, :, :, l, :, 0] = torch.floor(y).clamp(-2, h + 1).to(torch.int32)
        idx[:, :, :, l, :, 1] = torch.floor(x).clamp(-2, w + 1).to(torch.int32)
    return idx


def main():
    if not os.path.exists(REF_FILE):
        sys.exit("reference not present: golden vectors can only be regenerated in the build container")
    torch.manual_seed(0)
    torch.set_num_threads(1)
    ref = load_reference()
    for i, (name, cfg) in enumerate(CORE_CASES.items()):
        memory, loc, attn, grad_out = make_case_inputs(name, cfg, seed=1000 + i)
        res32 = run_reference_core(ref, memory, loc, attn, grad_out, cfg["H"], cfg["shapes"], torch.float32)
        res64 = run_reference_core(ref, memory, loc, attn, grad_out, cfg["H"], cfg["shapes"], torch.float64)
        out = dict(memory=memory, locations=loc, attention=attn, grad_out=grad_out,
                   shapes=np.asarray(cfg["shapes"], dtype=np.int32), n_heads=np.int32(cfg["H"]),
                   indices=reference_indices(loc, cfg["shapes"]))
        for tag, res in (("f32", res32), ("f64", res64)):
            for key, t in zip(("out", "grad_memory", "grad_locations", "grad_attention"), res):
                out[f"{key}_{tag}"] = t.to(torch.float32)
        np.savez_compressed(os.path.join(HERE, f"core_{name}.npz"),
                            **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in out.items()})
        print(f"core_{name}: out {tuple(res32[0].shape)} |out|max {res32[0].abs().max():.4f}")

    # module-level fixture: the reference MSDeformAttn with randomised parameters
    g = torch.Generator().manual_seed(77)
    d_model, n_levels, n_heads, n_points = 64, 3, 8, 4
    shapes = ((6, 8), (3, 4), (2, 2))
    N, nq, K = 2, 3, 5
    S = sum(h * w for h, w in shapes)
    mod = ref.MSDeformAttn(d_model=d_model, n_levels=n_levels, n_heads=n_heads, n_points=n_points)
    init_state = {k: v.detach().clone() for k, v in mod.state_dict().items()}
    with torch.no_grad():
        mod.sampling_offsets.weight.copy_(0.5 * torch.randn(mod.sampling_offsets.weight.shape, generator=g))
        mod.sampling_offsets.bias.add_(0.5 * torch.randn(mod.sampling_offsets.bias.shape, generator=g))
        mod.attention_weights.weight.copy_(torch.randn(mod.attention_weights.weight.shape, generator=g))
        mod.attention_weights.bias.copy_(torch.randn(mod.attention_weights.bias.shape, generator=g))
    query = torch.randn(N, nq * K, d_model, generator=g, requires_grad=True)
    ref_pts = torch.rand(N, nq, 1, K, 2, generator=g)
    memory = torch.randn(N, S, d_model, generator=g, requires_grad=True)
    grad_out = torch.randn(N, nq * K, d_model, generator=g)
    value = reference_value_list(memory, n_heads, shapes)
    out = mod(query, ref_pts, value, [list(s) for s in shapes])
    params = list(mod.parameters())
    grads = torch.autograd.grad(out, [query, memory, *params], grad_out)
    fixture = dict(query=query.detach(), reference_points=ref_pts, memory=memory.detach(), grad_out=grad_out,
                   shapes=np.asarray(shapes, dtype=np.int32), out=out.detach(),
                   grad_query=grads[0], grad_memory=grads[1],
                   hyper=np.asarray([d_model, n_levels, n_heads, n_points], dtype=np.int32))
    for (k, v) in mod.state_dict().items():
        fixture[f"param.{k}"] = v.detach()
    for (k, v) in init_state.items():
        fixture[f"init.{k}"] = v
    for (k, _), gparam in zip(mod.named_parameters(), grads[2:]):
        fixture[f"grad_param.{k}"] = gparam
    # reference init for the DETRPose-N style module (n_points % 4 != 0 -> zero bias, :311-312)
    mod_n = ref.MSDeformAttn(d_model=128, n_levels=2, n_heads=8, n_points=6)
    for (k, v) in mod_n.state_dict().items():
        fixture[f"init_n.{k}"] = v.detach().clone()
    np.savez_compressed(os.path.join(HERE, "module_small.npz"),
                        **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in fixture.items()})
    print("module_small: out", tuple(out.shape))


if __name__ == "__main__":
    main()

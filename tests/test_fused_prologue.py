"""Row f1: softmax + location prologue fused into the sampler, forward and backward.

Reference: ms_deform_attn.py:385-393 (Linear outputs, softmax over L*P), :412-416 (ref + offsets / (W_l, H_l)),
:440 -> :145-193 (core).  The fused launch is compared with (a) the reference's op sequence restated in
oracle/msda_torch.py on the same device (fp32, 1e-5 of max|ref| for the output and every gradient), (b) the
two-kernel path of this package (bit-identical output: same arithmetic), and (c) the golden module case of the
real reference class (tests/golden/module_small.npz).
"""
import numpy as np
import pytest
import torch

import detrpose_b200 as dp
from detrpose_b200 import functional as MF, synthetic, _lib
from oracle import msda_torch as otorch                       # checker only
from conftest import rel_err, load_module_case, value_list_from_memory

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5


def _inputs(workload="detrpose_s", N=2, Lq=70, seed=5, ref_levels=1):
    w = synthetic.WORKLOADS[workload]
    H, L, P = w["H"], len(w["shapes"]), w["P"]
    g = torch.Generator(device=DEV).manual_seed(seed)
    S = synthetic.pyramid_size(w["shapes"])
    memory = torch.randn(N, S, H * w["Dh"], device=DEV, generator=g)
    offsets = 2.5 * torch.randn(N, Lq, H * L * P * 2, device=DEV, generator=g)
    logits = torch.randn(N, Lq, H * L * P, device=DEV, generator=g)
    ref = torch.rand(N, Lq, ref_levels, 2, device=DEV, generator=g) * 1.1 - 0.05
    grad_out = torch.randn(N, Lq, H * w["Dh"], device=DEV, generator=g)
    return w, H, L, P, memory, offsets, logits, ref, grad_out


def _reference(w, H, L, P, memory, offsets, logits, ref, grad_out):
    """The reference's elementwise prologue + core through autograd (torch ops on the device)."""
    mem = memory.clone().requires_grad_(True)
    off = offsets.clone().requires_grad_(True)
    lg = logits.clone().requires_grad_(True)
    rf = ref.clone().requires_grad_(True)
    N, Lq = off.shape[:2]
    weights = torch.softmax(lg.view(N, Lq, H, L * P), -1).view(N, Lq, H, L, P)
    norm = torch.tensor([[wd, h] for h, wd in w["shapes"]], dtype=torch.float32, device=DEV).view(1, 1, 1, L, 1, 2)
    loc = rf[:, :, None, :, None, :] + off.view(N, Lq, H, L, P, 2) / norm
    out = otorch.core(otorch.make_value_list(mem, H, w["shapes"]), w["shapes"], loc, weights)
    grads = torch.autograd.grad(out, [mem, off, lg, rf], grad_out)
    return out.detach(), grads, loc.detach(), weights.detach()


@pytest.mark.parametrize("workload,ref_levels", [("detrpose_s", 1), ("detrpose_n", 1), ("detrpose_x", 1),
                                                 ("sweep4", 4)])
def test_fused_matches_reference_ops(workload, ref_levels):
    w, H, L, P, memory, offsets, logits, ref, grad_out = _inputs(workload, ref_levels=ref_levels)
    want_out, want_grads, _, _ = _reference(w, H, L, P, memory, offsets, logits, ref, grad_out)
    mem = memory.clone().requires_grad_(True)
    off = offsets.clone().requires_grad_(True)
    lg = logits.clone().requires_grad_(True)
    rf = ref.clone().requires_grad_(True)
    before = MF.stats["fused_forward_launches"]
    out = dp.functional.ms_deform_attn_fused(mem, w["shapes"], off, lg, rf, n_heads=H, n_levels=L, n_points=P)
    assert MF.stats["fused_forward_launches"] == before + 1          # the fused kernel did run
    grads = torch.autograd.grad(out, [mem, off, lg, rf], grad_out)
    assert rel_err(out.detach().cpu().numpy(), want_out.cpu().numpy()) <= TOL
    for name, g, wg in zip(("memory", "offsets", "logits", "ref"), grads, want_grads):
        assert g.shape == wg.shape, name
        assert rel_err(g.cpu().numpy(), wg.cpu().numpy()) <= TOL, name


def test_fused_output_is_bit_identical_to_the_two_kernel_path():
    w, H, L, P, memory, offsets, logits, ref, grad_out = _inputs(seed=9)
    with torch.no_grad():
        fused = dp.functional.ms_deform_attn_fused(memory, w["shapes"], offsets, logits, ref,
                                                   n_heads=H, n_levels=L, n_points=P)
        loc, att = dp.locations_and_weights(offsets, logits, ref, w["shapes"], H, L, P)
        two = dp.ms_deform_attn_core(memory, w["shapes"], loc, att)
    assert torch.equal(fused, two)


def test_fused_locations_keep_the_reference_indices():
    """The corner indices the fused kernels use are those of the reference's separately rounded divide + add:
    forward output equal to the core run on reference-computed locations (bitwise, fp32)."""
    w, H, L, P, memory, offsets, logits, ref, grad_out = _inputs(seed=13, Lq=200)
    _, _, loc, weights = _reference(w, H, L, P, memory, offsets, logits, ref, grad_out)
    with torch.no_grad():
        fused = dp.functional.ms_deform_attn_fused(memory, w["shapes"], offsets, logits, ref,
                                                   n_heads=H, n_levels=L, n_points=P)
        loc2, _ = dp.locations_and_weights(offsets, logits, ref, w["shapes"], H, L, P)
        core = dp.ms_deform_attn_core(memory, w["shapes"], loc, weights)
    assert torch.equal(loc2, loc)                                          # locations bit-exact
    assert rel_err(fused.cpu().numpy(), core.cpu().numpy()) <= 1e-6      # softmax differs by an ulp at most


def test_fused_bf16_value_under_autocast_like_inputs():
    w, H, L, P, memory, offsets, logits, ref, grad_out = _inputs(seed=21)
    mem_bf = memory.bfloat16()
    want_out, want_grads, _, _ = _reference(w, H, L, P, mem_bf.float(), offsets.bfloat16().float(),
                                            logits.bfloat16().float(), ref, grad_out.bfloat16().float())
    mem = mem_bf.clone().requires_grad_(True)
    off = offsets.bfloat16().requires_grad_(True)                          # Linear outputs under bf16 autocast
    lg = logits.bfloat16().requires_grad_(True)
    out = dp.functional.ms_deform_attn_fused(mem, w["shapes"], off, lg, ref, n_heads=H, n_levels=L, n_points=P)
    assert out.dtype == torch.bfloat16
    gm, go, gl = torch.autograd.grad(out, [mem, off, lg], grad_out.bfloat16())
    assert go.dtype == torch.bfloat16 and gl.dtype == torch.bfloat16 and gm.dtype == torch.bfloat16
    bf = 2.0 ** -8                                                         # one bf16 rounding of the result
    assert rel_err(out.detach().float().cpu().numpy(), want_out.cpu().numpy()) <= bf
    assert rel_err(gm.float().cpu().numpy(), want_grads[0].cpu().numpy()) <= bf
    assert rel_err(go.float().cpu().numpy(), want_grads[1].cpu().numpy()) <= bf
    assert rel_err(gl.float().cpu().numpy(), want_grads[2].cpu().numpy()) <= bf


def test_shared_value_across_fused_layers_accumulates_once():
    w, H, L, P, memory, _, _, _, _ = _inputs(seed=2)
    layers = [_inputs(seed=30 + i)[5:] for i in range(3)]
    mem = memory.clone().requires_grad_(True)
    value = value_list_from_memory(mem, H, w["shapes"])
    before = dict(MF.stats)
    loss = 0
    for off, lg, rf, go in layers:
        loss = loss + (dp.functional.ms_deform_attn_fused(value, w["shapes"], off, lg, rf,
                                                          n_heads=H, n_levels=L, n_points=P) * go).sum()
    loss.backward()
    d = {k: MF.stats[k] - before[k] for k in before}
    assert d["repack_launches"] == 1 and d["unpack_launches"] == 1 and d["backward_launches"] == 3, d
    ref_mem = memory.clone().requires_grad_(True)
    loss = 0
    for off, lg, rf, go in layers:
        N, Lq = off.shape[:2]
        weights = torch.softmax(lg.view(N, Lq, H, L * P), -1).view(N, Lq, H, L, P)
        norm = torch.tensor([[wd, h] for h, wd in w["shapes"]], dtype=torch.float32, device=DEV).view(1, 1, 1, L, 1, 2)
        loc = rf[:, :, None, :, None, :] + off.view(N, Lq, H, L, P, 2) / norm
        loss = loss + (otorch.core(otorch.make_value_list(ref_mem, H, w["shapes"]), w["shapes"], loc, weights) * go).sum()
    loss.backward()
    assert rel_err(mem.grad.cpu().numpy(), ref_mem.grad.cpu().numpy()) <= TOL


def test_patched_reference_module_forward_uses_the_fused_launch():
    from baseline import ref_harness as rh
    if not rh.available():
        pytest.skip("vendored reference (baseline/_ref) absent")
    ref = rh.load_reference()
    m = load_module_case()
    d_model, L, H, P = [int(v) for v in m["hyper"]]
    shapes = [[int(x) for x in s] for s in m["shapes"]]
    mod = ref.msda.MSDeformAttn(d_model=d_model, n_levels=L, n_heads=H, n_points=P).to(DEV)
    mod.load_state_dict({k[len("param."):]: torch.from_numpy(v) for k, v in m.items() if k.startswith("param.")})
    query = torch.from_numpy(m["query"]).to(DEV).requires_grad_(True)
    memory = torch.from_numpy(m["memory"]).to(DEV).requires_grad_(True)
    refp = torch.from_numpy(m["reference_points"]).to(DEV)
    dp.patch.install_forward(ref.msda)
    try:
        before = MF.stats["fused_forward_launches"]
        out = mod(query, refp, value_list_from_memory(memory, H, shapes), shapes)
        assert MF.stats["fused_forward_launches"] == before + 1
        gq, gm = torch.autograd.grad(out, [query, memory], torch.from_numpy(m["grad_out"]).to(DEV))
    finally:
        dp.patch.uninstall_forward(ref.msda)
    assert rel_err(out.detach().cpu().numpy(), m["out"]) <= TOL
    assert rel_err(gq.cpu().numpy(), m["grad_query"]) <= 2e-5
    assert rel_err(gm.cpu().numpy(), m["grad_memory"]) <= TOL

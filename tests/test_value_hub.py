"""Row f2: the value hand-over across the decoder layers that share one value list.

`transformer.py:1285-1286` builds the list once and `transformer.py:594-602` passes it to every decoder
layer.  The host side therefore repacks the reference's strided list ONCE, lets every layer's backward
launch add into one fp32 channel-last buffer (C ABI `accumulate=1`) and un-repacks ONCE; with the producer
patched (`patch.install_value_producer`) `memory` itself is read and neither copy happens.  Gradients are
compared with the reference's own op sequence (oracle/msda_torch.py) on the same device.
"""
import numpy as np
import pytest
import torch

import detrpose_b200 as dp
from detrpose_b200 import functional as MF, synthetic
from oracle import msda_torch as otorch                      # checker only
from conftest import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-5          # fp32, relative to max|ref| (BASELINE.json north_star)


def _stack_inputs(n_layers, N=2, Lq=90, workload="detrpose_s", seed=11):
    w = synthetic.WORKLOADS[workload]
    base = synthetic.make_inputs(N, Lq, w["H"], w["Dh"], w["shapes"], w["P"], seed=seed, device=DEV)
    layers = [synthetic.make_inputs(N, Lq, w["H"], w["Dh"], w["shapes"], w["P"], seed=seed + 1 + i, device=DEV)
              for i in range(n_layers)]
    return w, base["memory"], [(x["locations"], x["attention"], x["grad_out"]) for x in layers]


def _reference_grad(w, memory, layers, used=None):
    mem = memory.clone().requires_grad_(True)
    value = otorch.make_value_list(mem, w["H"], w["shapes"])
    loss = 0
    for i, (loc, att, go) in enumerate(layers):
        if used is not None and i not in used:
            continue
        loss = loss + (otorch.core(value, w["shapes"], loc, att) * go).sum()
    loss.backward()
    return mem.grad


def _snapshot():
    return dict(MF.stats)


def _delta(before):
    return {k: MF.stats[k] - before[k] for k in before}


def test_six_layers_one_repack_one_unrepack():
    w, memory, layers = _stack_inputs(6)
    mem = memory.clone().requires_grad_(True)
    value = otorch.make_value_list(mem, w["H"], w["shapes"])         # N = 2: the strided copy of the reference
    before = _snapshot()
    loss = 0
    for loc, att, go in layers:
        loss = loss + (dp.ms_deform_attn_core(value, w["shapes"], loc, att) * go).sum()
    loss.backward()
    d = _delta(before)
    assert d["repack_launches"] == 1 and d["unpack_launches"] == 1 and d["grad_handover"] == 1, d
    assert d["forward_launches"] == 6 and d["backward_launches"] == 6, d
    ref = _reference_grad(w, memory, layers)
    assert rel_err(mem.grad.cpu().numpy(), ref.cpu().numpy()) <= TOL


def test_patched_producer_reads_memory_zero_copy():
    w, memory, layers = _stack_inputs(4)
    mem = memory.clone().requires_grad_(True)
    value = MF.ValueList(mem, w["H"], [h * wd for h, wd in w["shapes"]])
    before = _snapshot()
    loss = 0
    for loc, att, go in layers:
        loss = loss + (dp.ms_deform_attn_core(value, w["shapes"], loc, att) * go).sum()
    loss.backward()
    d = _delta(before)
    assert d["repack_launches"] == 0 and d["unpack_launches"] == 0 and d["grad_handover"] == 1, d
    ref = _reference_grad(w, memory, layers)
    assert rel_err(mem.grad.cpu().numpy(), ref.cpu().numpy()) <= TOL
    # any other consumer still sees the reference's per-level tensors
    want = otorch.make_value_list(memory, w["H"], w["shapes"])
    assert len(value) == len(want) and all(torch.equal(a, b) for a, b in zip(value, want))


def test_layers_outside_the_graph_do_not_contribute():
    w, memory, layers = _stack_inputs(4)
    mem = memory.clone().requires_grad_(True)
    value = otorch.make_value_list(mem, w["H"], w["shapes"])
    outs = [dp.ms_deform_attn_core(value, w["shapes"], loc, att) for loc, att, _ in layers]
    used = {0, 2}
    loss = sum((outs[i] * layers[i][2]).sum() for i in used)
    loss.backward()
    ref = _reference_grad(w, memory, layers, used)
    assert rel_err(mem.grad.cpu().numpy(), ref.cpu().numpy()) <= TOL


def test_two_graphs_on_one_value_object():
    """Forward twice on the same list object, backward each graph on its own: no gradient leaks between them."""
    w, memory, layers = _stack_inputs(2)
    mem = memory.clone().requires_grad_(True)
    value = otorch.make_value_list(mem, w["H"], w["shapes"])
    a = (dp.ms_deform_attn_core(value, w["shapes"], layers[0][0], layers[0][1]) * layers[0][2]).sum()
    b = (dp.ms_deform_attn_core(value, w["shapes"], layers[1][0], layers[1][1]) * layers[1][2]).sum()
    (ga,) = torch.autograd.grad(a, mem, retain_graph=True)
    (gb,) = torch.autograd.grad(b, mem)
    ra = _reference_grad(w, memory, layers, {0})
    rb = _reference_grad(w, memory, layers, {1})
    assert rel_err(ga.cpu().numpy(), ra.cpu().numpy()) <= TOL
    assert rel_err(gb.cpu().numpy(), rb.cpu().numpy()) <= TOL


def test_training_loop_reuses_a_leaf_value():
    """The same leaf tensor over several iterations (new graph each time) keeps giving fresh gradients."""
    w, memory, layers = _stack_inputs(3)
    mem = memory.clone().requires_grad_(True)
    for it in range(3):
        mem.grad = None
        loc, att, go = layers[it]
        (dp.ms_deform_attn_core(mem, w["shapes"], loc, att) * go).sum().backward()
        ref = _reference_grad(w, memory, layers, {it})
        assert rel_err(mem.grad.cpu().numpy(), ref.cpu().numpy()) <= TOL


def test_inplace_edit_between_forward_and_backward_is_caught():
    w, memory, layers = _stack_inputs(1)
    mem = memory.clone().requires_grad_(True)
    loc, att, go = layers[0]
    loc = loc.clone().requires_grad_(True)
    scaled = mem * 1.0                                           # non-leaf, so it may be edited in place
    out = dp.ms_deform_attn_core(scaled, w["shapes"], loc, att)
    scaled.mul_(2.0)
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        (out * go).sum().backward()


def test_no_grad_forward_then_grad_forward_on_same_list():
    w, memory, layers = _stack_inputs(1)
    mem = memory.clone().requires_grad_(True)
    value = otorch.make_value_list(mem, w["H"], w["shapes"])
    loc, att, go = layers[0]
    with torch.no_grad():
        o0 = dp.ms_deform_attn_core(value, w["shapes"], loc, att)
    o1 = dp.ms_deform_attn_core(value, w["shapes"], loc, att)
    assert torch.equal(o0, o1) and o1.requires_grad
    (o1 * go).sum().backward()
    ref = _reference_grad(w, memory, layers)
    assert rel_err(mem.grad.cpu().numpy(), ref.cpu().numpy()) <= TOL


def test_no_device_memory_is_held_by_reference_cycles():
    """Pyramid, gradient buffer and token of a pass are freed with its graph (no wait for the cycle collector):
    the allocated device memory does not grow over repeated steps, with and without a backward pass."""
    import gc
    w, memory, layers = _stack_inputs(1, N=4, Lq=300)
    loc, att, go = layers[0]
    gc.collect()
    gc.disable()
    try:
        def step(backward):
            mem = memory.clone().requires_grad_(True)
            value = otorch.make_value_list(mem, w["H"], w["shapes"])          # N > 1: a repacked pyramid per step
            out = dp.ms_deform_attn_core(value, w["shapes"], loc, att)
            if backward:
                out.backward(go)
        for backward in (True, False):
            for _ in range(3):
                step(backward)
            torch.cuda.synchronize()
            base = torch.cuda.memory_allocated()
            for _ in range(10):
                step(backward)
            torch.cuda.synchronize()
            grown = torch.cuda.memory_allocated() - base
            assert grown <= memory.numel() * 4, (backward, grown)    # at most the one cached entry
    finally:
        gc.enable()

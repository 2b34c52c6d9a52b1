"""Parity of the sm_100a kernels (through the C ABI) with the reference.

Checkers: the committed golden vectors produced by the real reference
(tests/golden/, fp32 and fp64 runs), the numpy oracle (oracle/msda_numpy.py) as
the fp64 arbiter, and the reference's op sequence (oracle/msda_torch.py) on
seeded inputs.  Tolerances (BASELINE.json north_star):

* fp32: max|err| <= 1e-5 * max|ref| for the output and all three gradients;
* bf16 value/output mode: inputs are rounded to bf16 first and the comparison is made
  against the oracle evaluated on those rounded inputs; gradients (fp32) keep the 1e-5
  bound, the bf16 output is within one bf16 rounding: 2**-8 * max|ref|;
* sampling indices and level offsets: bit-exact.
"""
import threading

import numpy as np
import pytest
import torch

import detrpose_b200 as dp
from detrpose_b200 import synthetic, _lib
from conftest import core_case_names, load_core_case, load_module_case, rel_err, value_list_from_memory
from oracle import msda_numpy as onp
from oracle import msda_torch as otorch

pytestmark = pytest.mark.gpu
CASES = core_case_names()
TOL = 1e-5
TOL_BF16_OUT = 2.0 ** -8
DEV = "cuda:0"


def _run(case, layout, value_dtype=torch.float32):
    """Run forward+backward of the CUDA path for one golden case in the given value layout."""
    H, shapes = case["n_heads"], case["shapes"]
    mem = torch.from_numpy(case["memory"]).to(DEV).to(value_dtype).requires_grad_(True)
    loc = torch.from_numpy(case["locations"]).to(DEV).requires_grad_(True)
    att = torch.from_numpy(case["attention"]).to(DEV).requires_grad_(True)
    go = torch.from_numpy(case["grad_out"]).to(DEV).to(value_dtype)
    if layout == "reference":        # transformer.py:1285-1286: views (N=1) or spatial-innermost copy (N>1)
        value = value_list_from_memory(mem, H, shapes)
    elif layout == "contiguous":     # separately allocated contiguous per-level tensors
        value = [v.contiguous() for v in value_list_from_memory(mem, H, shapes)]
    elif layout == "memory":         # zero-copy (N, S, C)
        value = mem
    else:
        raise AssertionError(layout)
    # the golden vectors come from the reference on CPU, whose coordinate chain rounds every op
    out = dp.ms_deform_attn_core(value, shapes, loc, att, coord_mode=_lib.COORD_UNFUSED)
    gm, gl, ga = torch.autograd.grad(out, [mem, loc, att], go)
    torch.cuda.synchronize()
    return [t.detach().float().cpu().numpy() for t in (out, gm, gl, ga)]


def _arbiter(case, memory=None, grad_out=None):
    """fp64 oracle that takes the fp32 floor decisions (see oracle/msda_numpy.py:_corners)."""
    H, shapes = case["n_heads"], case["shapes"]
    memory = case["memory"] if memory is None else memory
    grad_out = case["grad_out"] if grad_out is None else grad_out
    N, S, C = memory.shape
    vals = [v.contiguous().numpy() for v in value_list_from_memory(torch.from_numpy(memory), H, shapes)]
    out = onp.msda_forward(vals, shapes, case["locations"], case["attention"], np.float64, np.float32)
    gv, gl, ga = onp.msda_backward(vals, shapes, case["locations"], case["attention"], grad_out,
                                   np.float64, np.float32)
    Dh = C // H
    gm = np.concatenate(gv, axis=2).reshape(N, H, Dh, S).transpose(0, 3, 1, 2).reshape(N, S, C)
    return out, gm, gl, ga


@pytest.fixture(params=[(1, 1), (0, 0), (2, 1), (3, 2), (107, 1)],
                ids=["lean+gather", "flat+flat", "staged+gather", "staged_small+gather512", "lean_plain+gather"])
def bwd_variant(request):
    """Run under every kernel variant.  Forward: 1 = lean (default), 0 = flat, 2 / 3 = TMA-staged coarse
    levels (one big CTA per SM / small CTAs), 107 = lean without its two defaults (one lane group per item
    also for rows of at most 32 bytes, no L2 prefetch); backward: 1 = gather form (default when the shape fits),
    2 = the same with 512 threads x 128 registers, 0 = flat + vector reductions.  A forced
    variant that does not support a shape fails loudly in the library; those combinations are skipped."""
    lib = _lib.load()
    lib.msda_b200_set_variant(*request.param)
    yield request.param
    lib.msda_b200_set_variant(-1, -1)


@pytest.mark.parametrize("layout", ["reference", "contiguous", "memory"])
@pytest.mark.parametrize("name", CASES)
def test_fp32_matches_reference_golden(name, layout, bwd_variant):
    c = load_core_case(name)
    out, gm, gl, ga = _run(c, layout)
    assert out.shape == c["out_f32"].shape
    assert rel_err(out, c["out_f32"]) <= TOL
    assert rel_err(gm, c["grad_memory_f32"]) <= TOL
    assert rel_err(gl, c["grad_locations_f32"]) <= TOL
    assert rel_err(ga, c["grad_attention_f32"]) <= TOL
    a_out, a_gm, a_gl, a_ga = _arbiter(c)
    assert rel_err(out, a_out) <= TOL
    assert rel_err(gm, a_gm) <= TOL
    assert rel_err(gl, a_gl) <= TOL
    assert rel_err(ga, a_ga) <= TOL


@pytest.mark.parametrize("name", CASES)
def test_bf16_value_mode(name, bwd_variant):
    c = load_core_case(name)
    mem_r = torch.from_numpy(c["memory"]).bfloat16().float().numpy()
    go_r = torch.from_numpy(c["grad_out"]).bfloat16().float().numpy()
    out, gm, gl, ga = _run(c, "reference", torch.bfloat16)
    a_out, a_gm, a_gl, a_ga = _arbiter(c, mem_r, go_r)
    assert rel_err(out, a_out) <= TOL_BF16_OUT
    assert rel_err(gl, a_gl) <= TOL
    assert rel_err(ga, a_ga) <= TOL
    assert rel_err(gm, a_gm) <= TOL_BF16_OUT       # grad w.r.t. a bf16 tensor is returned in bf16


@pytest.mark.parametrize("name", CASES)
def test_indices_and_level_offsets_bit_exact(name):
    c = load_core_case(name)
    loc = torch.from_numpy(c["locations"]).to(DEV)
    idx, starts = dp.sample_indices(loc, c["shapes"], coord_mode=_lib.COORD_UNFUSED)
    assert np.array_equal(idx.cpu().numpy(), c["indices"])
    assert starts.cpu().tolist() == onp.level_start_index(c["shapes"]).tolist()


def test_indices_bit_exact_at_scale_vs_torch_ops_on_device():
    """2e7 uniform samples per level size: the kernel's floor equals the reference's fp32 op chain
    (2*loc-1, +1, *size, -1, /2 as elementwise torch ops on the same device)."""
    g = torch.Generator(device=DEV).manual_seed(5)
    shapes = ((80, 80), (40, 40), (20, 20), (100, 75), (15, 25))
    loc = torch.rand((8, 20000, 8, len(shapes), 16, 2), device=DEV, generator=g) * 1.2 - 0.1
    idx, _ = dp.sample_indices(loc, shapes, coord_mode=_lib.COORD_UNFUSED)
    grid = 2 * loc - 1
    for l, (h, w) in enumerate(shapes):
        x = ((grid[:, :, :, l, :, 0] + 1) * w - 1) / 2
        y = ((grid[:, :, :, l, :, 1] + 1) * h - 1) / 2
        assert torch.equal(idx[:, :, :, l, :, 1], torch.floor(x).clamp(-2, w + 1).int())
        assert torch.equal(idx[:, :, :, l, :, 0], torch.floor(y).clamp(-2, h + 1).int())


def test_fma_chain_indices_bit_exact_at_scale():
    """COORD_FMA (the default: what ATen's CUDA sampler does) against the same chain with the fused
    multiply-subtract emulated exactly in fp64: (g+1)*size has at most 48 significant bits, so
    fp64 holds t*size - 1 exactly and one rounding to fp32 gives fma(t, size, -1)."""
    g = torch.Generator(device=DEV).manual_seed(6)
    shapes = ((80, 80), (40, 40), (20, 20), (100, 75), (15, 25))
    loc = torch.rand((8, 20000, 8, len(shapes), 16, 2), device=DEV, generator=g) * 1.2 - 0.1
    idx, _ = dp.sample_indices(loc, shapes, coord_mode=_lib.COORD_FMA)
    t = (2 * loc - 1) + 1
    for l, (h, w) in enumerate(shapes):
        x = ((t[:, :, :, l, :, 0].double() * w - 1).float()) / 2
        y = ((t[:, :, :, l, :, 1].double() * h - 1).float()) / 2
        assert torch.equal(idx[:, :, :, l, :, 1], torch.floor(x).clamp(-2, w + 1).int())
        assert torch.equal(idx[:, :, :, l, :, 0], torch.floor(y).clamp(-2, h + 1).int())
    assert dp.get_default_coord_mode() == _lib.COORD_FMA


@pytest.mark.parametrize("wl,N,Lq", [("detrpose_n", 1, 1080), ("detrpose_s", 2, 1080), ("detrpose_x", 1, 1080),
                                     ("detrpose_l", 2, 1476), ("sweep4", 2, 300)])
@pytest.mark.parametrize("degenerate", [False, True])
def test_model_shapes_vs_reference_ops_on_device(wl, N, Lq, degenerate, bwd_variant):
    """Real model shapes: CUDA kernels vs the reference's op sequence (F.grid_sample path) run on the
    same device, plus the fp64 arbiter for the forward."""
    w = synthetic.WORKLOADS[wl]
    inp = synthetic.make_inputs(N, Lq, w["H"], w["Dh"], w["shapes"], w["P"], seed=11, device=DEV,
                                degenerate=degenerate)
    mem = inp["memory"].requires_grad_(True)
    loc = inp["locations"].requires_grad_(True)
    att = inp["attention"].requires_grad_(True)
    value = otorch.make_value_list(mem, w["H"], w["shapes"])
    ref_out = otorch.core(value, w["shapes"], loc, att)
    ref_g = torch.autograd.grad(ref_out, [mem, loc, att], inp["grad_out"])
    out = dp.ms_deform_attn_core(otorch.make_value_list(mem, w["H"], w["shapes"]), w["shapes"], loc, att)
    got_g = torch.autograd.grad(out, [mem, loc, att], inp["grad_out"])
    assert rel_err(out.detach().cpu().numpy(), ref_out.detach().cpu().numpy()) <= TOL
    # grad_locations is discontinuous where a pixel coordinate is integral; samples whose floor
    # depends on whether "(g+1)*size - 1" is fused (ATen's CPU and CUDA builds differ there) are
    # excluded from that one comparison.  Output, grad_value and grad_attention are continuous.
    ia, _ = dp.sample_indices(loc.detach(), w["shapes"], coord_mode=_lib.COORD_UNFUSED)
    ib, _ = dp.sample_indices(loc.detach(), w["shapes"], coord_mode=_lib.COORD_FMA)
    keep = (ia == ib).all(-1, keepdim=True).float()
    assert keep.mean().item() > 0.9999
    assert rel_err(got_g[0].cpu().numpy(), ref_g[0].cpu().numpy()) <= TOL
    assert rel_err((got_g[1] * keep).cpu().numpy(), (ref_g[1] * keep).cpu().numpy()) <= TOL
    assert rel_err(got_g[2].cpu().numpy(), ref_g[2].cpu().numpy()) <= TOL


def test_random_shapes_against_reference_ops_on_device(bwd_variant):
    """Seeded sweep over odd shapes (tiny / non-square levels, 1..8 points, 1..4 levels, head dims 8..64,
    query counts that are not multiples of anything): every kernel variant against the reference's op
    sequence on the same device."""
    rng = np.random.default_rng(1234)
    for trial in range(24):
        N = int(rng.integers(1, 4)); H = int(rng.choice([1, 2, 4, 8])); Dh = int(rng.choice([8, 16, 24, 32, 48, 64]))
        L = int(rng.integers(1, 5)); P = int(rng.choice([1, 2, 3, 4, 6, 8])); Lq = int(rng.choice([1, 7, 33, 100, 257]))
        shapes = tuple((int(rng.integers(1, 25)), int(rng.integers(1, 25))) for _ in range(L))
        inp = synthetic.make_inputs(N, Lq, H, Dh, shapes, P, seed=100 + trial, device=DEV, clip=(-0.3, 1.3))
        mem = inp["memory"].requires_grad_(True)
        loc = inp["locations"].requires_grad_(True)
        att = inp["attention"].requires_grad_(True)
        ref_out = otorch.core(otorch.make_value_list(mem, H, shapes), shapes, loc, att)
        ref_g = torch.autograd.grad(ref_out, [mem, loc, att], inp["grad_out"])
        try:
            out = dp.ms_deform_attn_core(otorch.make_value_list(mem, H, shapes), shapes, loc, att)
            got_g = torch.autograd.grad(out, [mem, loc, att], inp["grad_out"])
        except _lib.MSDAError as exc:
            from conftest import MAY_REFUSE
            if MAY_REFUSE.get(tuple(bwd_variant), "\0") in str(exc):
                continue                                  # an opt-in variant refuses this shape (loudly)
            raise
        ia, _ = dp.sample_indices(loc.detach(), shapes, coord_mode=_lib.COORD_UNFUSED)
        ib, _ = dp.sample_indices(loc.detach(), shapes, coord_mode=_lib.COORD_FMA)
        keep = (ia == ib).all(-1, keepdim=True).float()
        tag = f"trial {trial}: N={N} H={H} Dh={Dh} L={L} P={P} Lq={Lq} shapes={shapes}"
        assert rel_err(out.detach().cpu().numpy(), ref_out.detach().cpu().numpy()) <= TOL, tag
        assert rel_err(got_g[0].cpu().numpy(), ref_g[0].cpu().numpy()) <= TOL, tag
        assert rel_err((got_g[1] * keep).cpu().numpy(), (ref_g[1] * keep).cpu().numpy()) <= TOL, tag
        assert rel_err(got_g[2].cpu().numpy(), ref_g[2].cpu().numpy()) <= TOL, tag


def test_full_size_properties():
    """BASELINE size (DETRPose-S, batch 64): properties that need no oracle.

    * adjoint identities: out is linear in value and in attention, so
      <out, g> == <value, grad_value> == <attention, grad_attention>;
    * linearity in value: out(a*v1 + b*v2) == a*out(v1) + b*out(v2);
    * a ones-pyramid with all samples inside the map returns sum(attention) == 1.
    """
    w = synthetic.WORKLOADS["detrpose_s"]
    N = 64
    inp = synthetic.make_inputs(N, w["Lq"], w["H"], w["Dh"], w["shapes"], w["P"], seed=21, device=DEV)
    mem = inp["memory"].requires_grad_(True)
    loc = inp["locations"].requires_grad_(True)
    att = inp["attention"].requires_grad_(True)
    g = inp["grad_out"]
    out = dp.ms_deform_attn_core(mem, w["shapes"], loc, att)
    gm, gl, ga = torch.autograd.grad(out, [mem, loc, att], g)
    lhs = (out.double() * g.double()).sum().item()
    assert abs((mem.double() * gm.double()).sum().item() - lhs) <= 1e-6 * abs(lhs) + 1e-3
    assert abs((att.double() * ga.double()).sum().item() - lhs) <= 1e-6 * abs(lhs) + 1e-3

    with torch.no_grad():
        mem2 = torch.randn_like(mem)
        o1 = dp.ms_deform_attn_core(mem.detach(), w["shapes"], loc, att)
        o2 = dp.ms_deform_attn_core(mem2, w["shapes"], loc, att)
        o12 = dp.ms_deform_attn_core(0.5 * mem.detach() - 2.0 * mem2, w["shapes"], loc, att)
        assert rel_err((0.5 * o1 - 2.0 * o2).cpu().numpy(), o12.cpu().numpy()) <= TOL
        inside = loc.detach().clamp(0.05, 0.95)
        ones = dp.ms_deform_attn_core(torch.ones_like(mem), w["shapes"], inside, att)
        assert (ones - 1.0).abs().max().item() <= 1e-5


@pytest.mark.parametrize("vdt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_backward_overwrite_and_accumulate_modes(vdt, bwd_variant):
    """DETRPose-S and -L (training length -> two query chunks) shapes.  Overwrite mode needs no zero-fill
    (the buffer starts as NaN); accumulate mode adds: two layers sharing a value give 2x the gradient.
    fp32 results are also held against the reference's op sequence on the device."""
    from detrpose_b200 import functional as MF
    for wl, N, Lq in (("detrpose_s", 2, 1080), ("detrpose_l", 1, 1800)):
        w = synthetic.WORKLOADS[wl]
        inp = synthetic.make_inputs(N, Lq, w["H"], w["Dh"], w["shapes"], w["P"], seed=4, device=DEV, value_dtype=vdt)
        pyr = MF.pack_value(inp["memory"], w["shapes"], w["H"])
        cm = MF.get_default_coord_mode()
        args = (pyr, inp["shapes"], inp["locations"], inp["attention"], inp["grad_out"], True, True, cm)
        gv1, gl1, ga1 = MF._backward_raw(*args)
        acc = torch.zeros_like(gv1)
        MF._backward_raw(*args, into=acc)
        MF._backward_raw(*args, into=acc)
        assert torch.isfinite(gv1).all()
        assert rel_err(acc.cpu().numpy(), 2.0 * gv1.cpu().numpy()) <= TOL
        if vdt is torch.float32:
            mem = inp["memory"].requires_grad_(True)
            loc = inp["locations"].requires_grad_(True)
            att = inp["attention"].requires_grad_(True)
            ref_out = otorch.core(otorch.make_value_list(mem, w["H"], w["shapes"]), w["shapes"], loc, att)
            rg = torch.autograd.grad(ref_out, [mem, loc, att], inp["grad_out"])
            ia, _ = dp.sample_indices(loc.detach(), w["shapes"], coord_mode=_lib.COORD_UNFUSED)
            ib, _ = dp.sample_indices(loc.detach(), w["shapes"], coord_mode=_lib.COORD_FMA)
            keep = (ia == ib).all(-1, keepdim=True).float()
            assert rel_err(gv1.reshape(rg[0].shape).cpu().numpy(), rg[0].cpu().numpy()) <= TOL
            assert rel_err((gl1 * keep).cpu().numpy(), (rg[1] * keep).cpu().numpy()) <= TOL
            assert rel_err(ga1.cpu().numpy(), rg[2].cpu().numpy()) <= TOL


def test_zero_attention_weights_and_masked_points(bwd_variant):
    """Exact zeros in the attention weights (masked points): grad_attention of those samples still
    needs the corner dots; grad_locations is exactly zero there."""
    c = load_core_case("s_like")
    att = c["attention"].copy()
    rng = np.random.default_rng(0)
    mask = rng.random(att.shape) < 0.3
    att[mask] = 0.0
    c2 = dict(c, attention=att)
    out, gm, gl, ga = _run(c2, "reference")
    a_out, a_gm, a_gl, a_ga = _arbiter(c2)
    assert rel_err(out, a_out) <= TOL and rel_err(gm, a_gm) <= TOL
    assert rel_err(gl, a_gl) <= TOL and rel_err(ga, a_ga) <= TOL
    assert np.all(gl[mask] == 0.0)
    assert np.abs(ga[mask]).max() > 0.0


def test_repack_cache_never_serves_stale_values():
    w = synthetic.WORKLOADS["detrpose_n"]
    inp = synthetic.make_inputs(2, 36, w["H"], w["Dh"], w["shapes"], w["P"], seed=3, device=DEV)
    from detrpose_b200 import functional as MF
    value = otorch.make_value_list(inp["memory"], w["H"], w["shapes"])
    before = MF.stats["repack_launches"]
    a = dp.ms_deform_attn_core(value, w["shapes"], inp["locations"], inp["attention"])
    b = dp.ms_deform_attn_core(value, w["shapes"], inp["locations"], inp["attention"])   # cache hit
    assert MF.stats["repack_launches"] == before + 1      # decoder layers sharing one list repack once
    assert torch.equal(a, b)
    value[0].mul_(2.0)                                   # in-place edit bumps the version counter
    c = dp.ms_deform_attn_core(value, w["shapes"], inp["locations"], inp["attention"])
    ref = otorch.core(value, w["shapes"], inp["locations"], inp["attention"])
    assert rel_err(c.cpu().numpy(), ref.cpu().numpy()) <= TOL
    assert not torch.equal(a, c)


def test_module_matches_reference_module_golden():
    m = load_module_case()
    d_model, L, H, P = [int(v) for v in m["hyper"]]
    shapes = [tuple(int(x) for x in s) for s in m["shapes"]]
    mod = dp.MSDeformAttn(d_model=d_model, n_levels=L, n_heads=H, n_points=P).to(DEV)
    mod.load_state_dict({k[len("param."):]: torch.from_numpy(v) for k, v in m.items() if k.startswith("param.")})
    query = torch.from_numpy(m["query"]).to(DEV).requires_grad_(True)
    memory = torch.from_numpy(m["memory"]).to(DEV).requires_grad_(True)
    refp = torch.from_numpy(m["reference_points"]).to(DEV)
    value = value_list_from_memory(memory, H, shapes)
    out = mod(query, refp, value, [list(s) for s in shapes])
    params = dict(mod.named_parameters())
    grads = torch.autograd.grad(out, [query, memory, *params.values()], torch.from_numpy(m["grad_out"]).to(DEV))
    assert rel_err(out.detach().cpu().numpy(), m["out"]) <= TOL
    assert rel_err(grads[0].cpu().numpy(), m["grad_query"]) <= 2e-5       # two GEMMs deep
    assert rel_err(grads[1].cpu().numpy(), m["grad_memory"]) <= TOL
    for (k, _), gk in zip(params.items(), grads[2:]):
        assert rel_err(gk.cpu().numpy(), m[f"grad_param.{k}"]) <= 2e-5, k
    # inference path (fused prologue kernel) gives the same output
    with torch.no_grad():
        out_ng = mod(query.detach(), refp, value_list_from_memory(memory.detach(), H, shapes), shapes)
    assert rel_err(out_ng.cpu().numpy(), m["out"]) <= TOL


def test_fused_prologue_locations_bit_exact():
    m = load_module_case()
    d_model, L, H, P = [int(v) for v in m["hyper"]]
    shapes = [tuple(int(x) for x in s) for s in m["shapes"]]
    g = torch.Generator(device=DEV).manual_seed(9)
    offsets = torch.randn((3, 50, H * L * P * 2), device=DEV, generator=g) * 3
    logits = torch.randn((3, 50, H * L * P), device=DEV, generator=g) * 2
    ref = torch.rand((3, 50, 1, 2), device=DEV, generator=g)
    loc, att = dp.locations_and_weights(offsets, logits, ref, shapes, H, L, P)
    norm = torch.tensor(shapes, device=DEV).flip([1]).reshape(1, 1, 1, L, 1, 2)
    loc_ref = ref[:, :, None, :, None, :] + offsets.view(3, 50, H, L, P, 2) / norm
    att_ref = torch.softmax(logits.view(3, 50, H, L * P), -1).view(3, 50, H, L, P)
    assert torch.equal(loc, loc_ref)
    assert rel_err(att.cpu().numpy(), att_ref.cpu().numpy()) <= 1e-6
    # analytic backward of the fused prologue against autograd through the reference's torch ops
    o1, l1, r1 = (t.clone().requires_grad_(True) for t in (offsets, logits, ref))
    o2, l2, r2 = (t.clone().requires_grad_(True) for t in (offsets, logits, ref))
    gl, ga = torch.randn_like(loc), torch.randn_like(att)
    loc1, att1 = dp.locations_and_weights(o1, l1, r1, shapes, H, L, P)
    g1 = torch.autograd.grad([loc1, att1], [o1, l1, r1], [gl, ga])
    loc2 = r2[:, :, None, :, None, :] + o2.view(3, 50, H, L, P, 2) / norm
    att2 = torch.softmax(l2.view(3, 50, H, L * P), -1).view(3, 50, H, L, P)
    g2 = torch.autograd.grad([loc2, att2], [o2, l2, r2], [gl, ga])
    for a, b in zip(g1, g2):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) <= 2e-6


def test_backward_from_worker_thread_and_non_default_stream():
    c = load_core_case("s_like")
    result = {}

    def work():
        s = torch.cuda.Stream(device=DEV)
        with torch.cuda.stream(s):
            result["v"] = _run(c, "reference")
        s.synchronize()

    t = threading.Thread(target=work)
    t.start()
    t.join()
    out, gm, gl, ga = result["v"]
    assert rel_err(out, c["out_f32"]) <= TOL and rel_err(gm, c["grad_memory_f32"]) <= TOL


def test_forward_backward_capture_in_cuda_graph():
    """The C-ABI calls only enqueue work on the given stream (no allocation, no synchronisation), so a
    forward + backward pair can be captured once and replayed on new data in the same buffers."""
    from detrpose_b200 import functional as MF
    w = synthetic.WORKLOADS["detrpose_s"]
    mode = MF.get_default_coord_mode()

    def inputs(seed):
        return synthetic.make_inputs(3, 200, w["H"], w["Dh"], w["shapes"], w["P"], seed=seed, device=DEV,
                                     value_dtype=torch.bfloat16)
    a = inputs(1)
    pyr = MF.pack_value(a["memory"], w["shapes"], w["H"]).clone()
    loc, attn, go = a["locations"].clone(), a["attention"].clone(), a["grad_out"].clone()
    s = torch.cuda.Stream(device=DEV)
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):                                   # warm-up outside the capture
            out = MF._forward_raw(pyr, w["shapes"], loc, attn, torch.bfloat16, mode)
            gv, gl, ga = MF._backward_raw(pyr, w["shapes"], loc, attn, go, True, True, mode)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        out = MF._forward_raw(pyr, w["shapes"], loc, attn, torch.bfloat16, mode)
        gv, gl, ga = MF._backward_raw(pyr, w["shapes"], loc, attn, go, True, True, mode)
    b = inputs(2)
    pyr.copy_(MF.pack_value(b["memory"], w["shapes"], w["H"]))
    loc.copy_(b["locations"]); attn.copy_(b["attention"]); go.copy_(b["grad_out"])
    graph.replay()
    torch.cuda.synchronize()
    want_out = MF._forward_raw(pyr, w["shapes"], loc, attn, torch.bfloat16, mode)
    want = MF._backward_raw(pyr, w["shapes"], loc, attn, go, True, True, mode)
    assert torch.equal(out, want_out)
    for x, y in zip((gv, gl, ga), want):
        assert rel_err(x.cpu().numpy(), y.cpu().numpy()) <= 1e-6


def test_fp16_inputs_run_the_sampler_in_fp32():
    """The reference trains under fp16 autocast (engine.py:20), where grid_sample runs in fp32: fp16 value
    lists / tensors are up-cast, the result is fp32 and matches the fp32 run on the same (fp16-rounded) data."""
    c = load_core_case("s_like")
    mem16 = torch.from_numpy(c["memory"]).to(DEV).half()
    loc = torch.from_numpy(c["locations"]).to(DEV)
    att = torch.from_numpy(c["attention"]).to(DEV)
    want = dp.ms_deform_attn_core(mem16.float(), c["shapes"], loc, att, n_heads=c["n_heads"])
    got_t = dp.ms_deform_attn_core(mem16, c["shapes"], loc.half(), att.half(), n_heads=c["n_heads"])
    got_l = dp.ms_deform_attn_core(value_list_from_memory(mem16, c["n_heads"], c["shapes"]), c["shapes"], loc, att)
    assert got_t.dtype == torch.float32 and got_l.dtype == torch.float32
    assert torch.equal(got_l, want)
    assert rel_err(got_t.cpu().numpy(), want.cpu().numpy()) < 5e-3       # locations / attention rounded to fp16
    with pytest.raises(TypeError):
        dp.ms_deform_attn_core(value_list_from_memory(mem16.double(), c["n_heads"], c["shapes"]), c["shapes"], loc, att)


def test_error_behaviour_on_device():
    loc = torch.rand(1, 2, 2, 1, 2, 2, device=DEV)
    att = torch.full((1, 2, 2, 1, 2), 0.5, device=DEV)
    with pytest.raises(_lib.MSDAError, match="Dh=4"):
        dp.ms_deform_attn_core([torch.randn(2, 4, 16, device=DEV)], [(4, 4)], loc, att)
    with pytest.raises(ValueError):
        dp.ms_deform_attn_core([torch.randn(2, 8, 15, device=DEV)], [(4, 4)], loc, att)
    with pytest.raises(TypeError):
        dp.ms_deform_attn_core(torch.randn(1, 16, 16, device=DEV, dtype=torch.float64), [(4, 4)],
                               loc, att, n_heads=2)


def test_module_under_bf16_autocast_uses_fused_prologue_and_tracks_fp32():
    """bf16 autocast (Linear outputs in bf16): the module still takes the fused fp32 prologue, returns the
    value dtype, back-propagates to query / value / parameters, and stays close to its own fp32 run."""
    torch.manual_seed(0)
    m = dp.MSDeformAttn(d_model=256, n_levels=3, n_heads=8, n_points=4).to(DEV)
    with torch.no_grad():
        m.sampling_offsets.weight.normal_(0, 0.02)
        m.attention_weights.weight.normal_(0, 0.05)
    shapes = [[20, 20], [10, 10], [5, 5]]
    S = sum(h * w for h, w in shapes)
    query = torch.randn(2, 36, 256, device=DEV, requires_grad=True)
    ref = torch.rand(2, 2, 1, 18, 2, device=DEV)
    memory = torch.randn(2, S, 256, device=DEV, requires_grad=True)
    want = m(query, ref, memory, shapes)
    before = dp.functional.stats["forward_launches"]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        got = m(query, ref, memory.bfloat16(), shapes)
    assert dp.functional.stats["forward_launches"] == before + 1
    assert got.dtype == torch.bfloat16
    assert rel_err(got.float().detach().cpu().numpy(), want.detach().cpu().numpy()) < 0.05
    got.float().square().mean().backward()
    assert query.grad is not None and memory.grad is not None and m.sampling_offsets.weight.grad is not None
    assert torch.isfinite(query.grad).all() and torch.isfinite(memory.grad).all()


# --------------------------------------------------------------------------------------------------
# Round-2 additions: the bench configuration itself, the training-time pyramids, the big sweep point,
# tiny attention weights (VERDICT r01 "What's weak" 1-4)
# --------------------------------------------------------------------------------------------------
def _device_reference(w_shapes, H, mem, loc, att, go):
    """The reference's op sequence on the device (fp32), value handed over as its own strided list."""
    m = mem.detach().float().requires_grad_(True)
    l = loc.detach().requires_grad_(True)
    a = att.detach().requires_grad_(True)
    out = otorch.core(otorch.make_value_list(m, H, w_shapes), w_shapes, l, a)
    g = torch.autograd.grad(out, [m, l, a], go.float())
    return out.detach(), g


def _floor_agree(loc, shapes):
    ia, _ = dp.sample_indices(loc.detach(), shapes, coord_mode=_lib.COORD_UNFUSED)
    ib, _ = dp.sample_indices(loc.detach(), shapes, coord_mode=_lib.COORD_FMA)
    return (ia == ib).all(-1, keepdim=True).float()


@pytest.mark.parametrize("layout", ["memory", "reference"])
@pytest.mark.parametrize("vdt", [torch.bfloat16, torch.float32], ids=["bf16", "f32"])
def test_bench_configuration_against_reference_ops(vdt, layout):
    """Exactly what bench.py times -- DETRPose-S shape, batch 64, bf16 (and fp32) storage, zero-copy memory and
    the reference's strided list -- held against the reference's ops on the same device."""
    w = synthetic.WORKLOADS["detrpose_s"]
    N = 64
    inp = synthetic.make_inputs(N, w["Lq"], w["H"], w["Dh"], w["shapes"], w["P"], seed=0, device=DEV, value_dtype=vdt)
    ref_out, ref_g = _device_reference(w["shapes"], w["H"], inp["memory"], inp["locations"], inp["attention"],
                                       inp["grad_out"])
    mem = inp["memory"].clone().requires_grad_(True)
    loc = inp["locations"].clone().requires_grad_(True)
    att = inp["attention"].clone().requires_grad_(True)
    value = mem if layout == "memory" else otorch.make_value_list(mem, w["H"], w["shapes"])
    out = dp.ms_deform_attn_core(value, w["shapes"], loc, att)
    gm, gl, ga = torch.autograd.grad(out, [mem, loc, att], inp["grad_out"])
    keep = _floor_agree(loc, w["shapes"])
    # bf16 storage: output and the returned value gradient carry one bf16 rounding; the fp32 results 1e-5
    tol_store = 2.0 ** -8 if vdt == torch.bfloat16 else TOL
    assert rel_err(out.detach().float().cpu().numpy(), ref_out.cpu().numpy()) <= tol_store
    assert rel_err(gm.float().cpu().numpy(), ref_g[0].cpu().numpy()) <= tol_store
    assert rel_err((gl * keep).cpu().numpy(), (ref_g[1] * keep).cpu().numpy()) <= TOL
    assert rel_err(ga.cpu().numpy(), ref_g[2].cpu().numpy()) <= TOL


@pytest.mark.parametrize("shapes", [((100, 100), (50, 50), (25, 25)), ((60, 60), (30, 30), (15, 15)),
                                    ((100, 75), (50, 38), (25, 19))], ids=["800px", "480px", "800x600px"])
@pytest.mark.parametrize("Lq", [1476, 1584, 1800])
@pytest.mark.parametrize("vdt", [torch.float32, torch.bfloat16], ids=["f32", "bf16"])
def test_training_pyramids_and_denoising_lengths(shapes, Lq, vdt):
    """Multi-scale collate gives 480..800 px inputs (src/data/dataloader.py:56-61,103-105) and the denoising
    queries Len_q = (pad + 60) * 18 (dn_component.py:54-63): several query chunks, 101 x 101 bins."""
    H, Dh, P, N = 8, 32, 4, 2
    inp = synthetic.make_inputs(N, Lq, H, Dh, shapes, P, seed=Lq, device=DEV, value_dtype=vdt)
    ref_out, ref_g = _device_reference(shapes, H, inp["memory"], inp["locations"], inp["attention"], inp["grad_out"])
    mem = inp["memory"].clone().requires_grad_(True)
    loc = inp["locations"].clone().requires_grad_(True)
    att = inp["attention"].clone().requires_grad_(True)
    out = dp.ms_deform_attn_core(mem, shapes, loc, att)
    gm, gl, ga = torch.autograd.grad(out, [mem, loc, att], inp["grad_out"])
    keep = _floor_agree(loc, shapes)
    tol_store = 2.0 ** -8 if vdt == torch.bfloat16 else TOL
    assert rel_err(out.detach().float().cpu().numpy(), ref_out.cpu().numpy()) <= tol_store
    assert rel_err(gm.float().cpu().numpy(), ref_g[0].cpu().numpy()) <= tol_store
    assert rel_err((gl * keep).cpu().numpy(), (ref_g[1] * keep).cpu().numpy()) <= TOL
    assert rel_err(ga.cpu().numpy(), ref_g[2].cpu().numpy()) <= TOL


def test_sweep4_batch32_lq3000():
    """BASELINE configs[4] corner: 4 levels, Len_q 3000 (three query chunks), batch 32, fp32."""
    w = synthetic.WORKLOADS["sweep4"]
    N, Lq = 32, 3000
    inp = synthetic.make_inputs(N, Lq, w["H"], w["Dh"], w["shapes"], w["P"], seed=4, device=DEV)
    ref_out, ref_g = _device_reference(w["shapes"], w["H"], inp["memory"], inp["locations"], inp["attention"],
                                       inp["grad_out"])
    mem = inp["memory"].clone().requires_grad_(True)
    loc = inp["locations"].clone().requires_grad_(True)
    att = inp["attention"].clone().requires_grad_(True)
    out = dp.ms_deform_attn_core(otorch.make_value_list(mem, w["H"], w["shapes"]), w["shapes"], loc, att)
    gm, gl, ga = torch.autograd.grad(out, [mem, loc, att], inp["grad_out"])
    keep = _floor_agree(loc, w["shapes"])
    assert rel_err(out.detach().cpu().numpy(), ref_out.cpu().numpy()) <= TOL
    assert rel_err(gm.cpu().numpy(), ref_g[0].cpu().numpy()) <= TOL
    assert rel_err((gl * keep).cpu().numpy(), (ref_g[1] * keep).cpu().numpy()) <= TOL
    assert rel_err(ga.cpu().numpy(), ref_g[2].cpu().numpy()) <= TOL


@pytest.mark.parametrize("tiny", [1e-20, 1e-31, 0.0, -1e-31])
def test_tiny_attention_weights(tiny, bwd_variant):
    """Samples whose weight is below 1e-30 in magnitude are finished outside the pixel pass of the gather
    backward (msda_bwd_gather.cu, P1): their contribution to grad_value / grad_locations (< 1e-30 |grad_out|)
    is dropped, grad_attention is still exact.  1e-20 takes the normal path.  Either way every gradient
    stays within 1e-5 of max|ref|, and nothing becomes NaN (the weight is recovered by a division)."""
    c = load_core_case("s_like")
    att = c["attention"].copy()
    rng = np.random.default_rng(7)
    mask = rng.random(att.shape) < 0.25
    att[mask] = tiny
    c2 = dict(c, attention=att)
    out, gm, gl, ga = _run(c2, "reference")
    a_out, a_gm, a_gl, a_ga = _arbiter(c2)
    for got in (out, gm, gl, ga):
        assert np.isfinite(got).all()
    assert rel_err(out, a_out) <= TOL and rel_err(gm, a_gm) <= TOL
    assert rel_err(gl, a_gl) <= TOL and rel_err(ga, a_ga) <= TOL
    # absolute worst case of the dropped contributions: below 1e-30 * |grad_out| * |value| per sample
    if abs(tiny) < 1e-30:
        assert np.abs(gl[mask] - a_gl[mask]).max() <= 1e-25


@pytest.mark.parametrize("Dh,vdt", [(8, torch.bfloat16), (16, torch.bfloat16), (8, torch.float32), (16, torch.float32),
                                    (32, torch.bfloat16)])
def test_forward_forms_agree(Dh, vdt):
    """The forms the lean forward chooses between (msda_fwd.cu): two lane groups per item for channel rows of
    at most 32 bytes (the default there) against one (variant 107) and against the flat kernel (0), on an
    item count that leaves the last CTA partly filled; and the L2 prefetch (a hint) changes no bit."""
    from detrpose_b200 import functional as MF
    lib = _lib.load()
    H, P, shapes, N, Lq = 4, 6, ((13, 17), (7, 9)), 3, 77              # 924 items: not a multiple of 32 .. 256
    inp = synthetic.make_inputs(N, Lq, H, Dh, shapes, P, seed=5, device=DEV, clip=(-0.3, 1.3), value_dtype=vdt)
    pyr = MF.pack_value(inp["memory"], inp["shapes"], H)
    loc, att = inp["locations"], inp["attention"]
    cm = MF.get_default_coord_mode()
    outs = {}
    try:
        for v in (-1, 150, 107, 0):
            lib.msda_b200_set_variant(v, -1)
            outs[v] = MF._forward_raw(pyr, inp["shapes"], loc, att, torch.float32, cm)
    finally:
        lib.msda_b200_set_variant(-1, -1)
    torch.cuda.synchronize()
    assert torch.equal(outs[-1], outs[150])
    ref = otorch.core(otorch.make_value_list(inp["memory"].float(), H, shapes), shapes, loc, att)
    for v in (-1, 107, 0):
        assert rel_err(outs[v].cpu().numpy(), ref.cpu().numpy()) <= TOL, v
    if Dh * (2 if vdt == torch.bfloat16 else 4) > 32:
        assert torch.equal(outs[-1], outs[107])            # one form only for wider rows

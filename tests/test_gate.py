"""Gate epilogue: oracle vs the real reference's golden vectors (CPU) and the
sm_100a kernels vs both (GPU, through the C ABI).

Golden vectors: tests/golden/gate_*.npz from tests/golden/make_golden_blocks.py (the reference's own
``Gate`` class, transformer.py:222-235, fp32 and fp64 runs).  Tolerances: fp32 within 1e-5 of max|ref|
for the output and every gradient (2e-5 for the row-summed parameter gradients of the kernel, whose
fp32 atomics have no fixed order); bf16 I/O within one bf16 rounding (2^-8 of max|ref|).
"""
import glob
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, rel_err
from oracle import gate_numpy as og

_spec = importlib.util.spec_from_file_location("make_golden_blocks", os.path.join(GOLDEN_DIR, "make_golden_blocks.py"))
_mgb = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_mgb)

CASES = sorted(os.path.basename(p)[len("gate_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "gate_*.npz")))


def load_gate_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"gate_{name}.npz"))
    c = {k: z[k] for k in z.files}
    C = c["x1"].shape[-1]
    if "param_gate_weight" not in c:
        c["param_gate_weight"] = _mgb.structured_weight(C)
    c["C"] = C
    return c


def _check_weight_grad(c, gw, tag, tol):
    if f"grad_gate_weight_{tag}" in c:
        assert rel_err(gw, c[f"grad_gate_weight_{tag}"]) < tol
    else:
        left, right = _mgb.probe_vectors(c["C"])
        assert rel_err(np.asarray(gw, np.float64) @ right.astype(np.float64), c[f"grad_gate_weight_right_{tag}"]) < tol
        assert rel_err(left.astype(np.float64) @ np.asarray(gw, np.float64), c[f"grad_gate_weight_left_{tag}"]) < tol


def test_gate_golden_present():
    assert len(CASES) >= 3


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("dtype,tag,tol", [(np.float32, "f32", 1e-5), (np.float64, "f64", 1e-6)])
def test_oracle_gate_forward_backward(name, dtype, tag, tol):
    c = load_gate_case(name)
    args = (c["x1"], c["x2"], c["param_gate_weight"], c["param_gate_bias"], c["param_norm_weight"], c["param_norm_bias"])
    y = og.gate_forward(*args, eps=float(c["eps"]), dtype=dtype)
    assert rel_err(y, c[f"y_{tag}"]) < tol
    g = og.gate_backward(*args, c["grad_y"], eps=float(c["eps"]), dtype=dtype)
    assert rel_err(g["x1"], c[f"grad_x1_{tag}"]) < tol
    assert rel_err(g["x2"], c[f"grad_x2_{tag}"]) < tol
    assert rel_err(g["bias"], c[f"grad_gate_bias_{tag}"]) < tol
    assert rel_err(g["gamma"], c[f"grad_norm_weight_{tag}"]) < tol
    assert rel_err(g["beta"], c[f"grad_norm_bias_{tag}"]) < tol
    _check_weight_grad(c, g["weight"], tag, tol)


def test_reference_init_recorded():
    """transformer.py:226-228: zero gate weight, bias -log((1 - 0.5) / 0.5) = 0."""
    c = load_gate_case(CASES[0])
    assert float(c["init_gate_weight_absmax"]) == 0.0
    assert np.all(c["init_gate_bias"] == 0.0)


def test_gate_module_init_and_names():
    from detrpose_b200.gate import Gate
    g = Gate(256)
    assert sorted(g.state_dict()) == ["gate.bias", "gate.weight", "norm.bias", "norm.weight"]
    assert g.gate.weight.shape == (512, 512) and float(g.gate.weight.detach().abs().max()) == 0.0
    assert float(g.gate.bias.detach().abs().max()) == 0.0
    with pytest.raises(ValueError):
        Gate(100)
    with pytest.raises(RuntimeError, match="no CPU path"):
        g(torch.zeros(2, 256), torch.zeros(2, 256))


# ----------------------------------------------------------------------------------------------
# GPU: kernels through the C ABI
# ----------------------------------------------------------------------------------------------
def _module_from_case(c, device, dtype=torch.float32):
    from detrpose_b200.gate import Gate
    m = Gate(c["C"])
    m.load_state_dict({"gate.weight": torch.from_numpy(c["param_gate_weight"]),
                       "gate.bias": torch.from_numpy(c["param_gate_bias"]),
                       "norm.weight": torch.from_numpy(c["param_norm_weight"]),
                       "norm.bias": torch.from_numpy(c["param_norm_bias"])})
    return m.to(device=device, dtype=dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_gate_module_golden(name):
    c = load_gate_case(name)
    dev = "cuda:0"
    m = _module_from_case(c, dev)
    x1 = torch.from_numpy(c["x1"]).to(dev).requires_grad_(True)
    x2 = torch.from_numpy(c["x2"]).to(dev).requires_grad_(True)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        y = m(x1, x2)
        grads = torch.autograd.grad(y, [x1, x2, *m.parameters()], torch.from_numpy(c["grad_y"]).to(dev))
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    names = [n for n, _ in m.named_parameters()]
    for tag in ("f32", "f64"):
        assert rel_err(y.detach().cpu().numpy(), c[f"y_{tag}"]) < 1e-5
        assert rel_err(grads[0].cpu().numpy(), c[f"grad_x1_{tag}"]) < 1e-5
        assert rel_err(grads[1].cpu().numpy(), c[f"grad_x2_{tag}"]) < 1e-5
        for n, g in zip(names, grads[2:]):
            if n == "gate.weight":
                _check_weight_grad(c, g.cpu().numpy(), tag, 2e-5)
            else:
                assert rel_err(g.cpu().numpy(), c[f"grad_{n.replace('.', '_')}_{tag}"]) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("C", [128, 256, 384, 512])
@pytest.mark.parametrize("rows", [1, 37, 4099])
def test_gpu_gate_epilogue_vs_oracle(C, rows):
    """Kernel boundary itself (pre-activations in, no GEMM): against the fp64 oracle and the torch ops."""
    from detrpose_b200.gate import gate_epilogue
    g = torch.Generator().manual_seed(C + rows)
    dev = "cuda:0"
    pre = (2.0 * torch.randn(rows, 2 * C, generator=g)).to(dev).requires_grad_(True)
    x1 = torch.randn(rows, C, generator=g).to(dev).requires_grad_(True)
    x2 = (3.0 * torch.randn(rows, C, generator=g) + 0.5).to(dev).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(dev).requires_grad_(True)
    beta = (0.2 * torch.randn(C, generator=g)).to(dev).requires_grad_(True)
    gy = torch.randn(rows, C, generator=g).to(dev)
    y = gate_epilogue(pre, x1, x2, gamma, beta, 1e-5)
    grads = torch.autograd.grad(y, [pre, x1, x2, gamma, beta], gy)

    # fp64 arbiter: the oracle's formulas through torch autograd in double on the CPU
    p64, a64, b64, g64, bt64 = (t.detach().double().cpu().requires_grad_(True) for t in (pre, x1, x2, gamma, beta))
    y64, _, _ = og.gate_epilogue_forward(p64.detach().numpy(), a64.detach().numpy(), b64.detach().numpy(),
                                         g64.detach().numpy(), bt64.detach().numpy(), 1e-5)
    assert rel_err(y.detach().cpu().numpy(), y64) < 1e-5
    s = torch.sigmoid(p64)
    z = s[:, :C] * a64 + s[:, C:] * b64
    ref = torch.nn.functional.layer_norm(z, (C,), g64, bt64, 1e-5)
    rg = torch.autograd.grad(ref, [p64, a64, b64, g64, bt64], gy.double().cpu())
    for ours, want, tol in zip(grads, rg, (1e-5, 1e-5, 1e-5, 2e-5, 2e-5)):
        assert rel_err(ours.cpu().numpy(), want.numpy()) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("pre_dtype,x_dtype", [(torch.bfloat16, torch.bfloat16), (torch.bfloat16, torch.float32),
                                               (torch.float32, torch.bfloat16)])
def test_gpu_gate_epilogue_bf16(pre_dtype, x_dtype):
    """bf16 I/O, fp32 arithmetic: compare with the fp64 formulas evaluated on the bf16-rounded inputs;
    tolerance one bf16 rounding of the result (2^-8 of max|ref|) for bf16 outputs, 1e-5 for fp32 ones."""
    from detrpose_b200.gate import gate_epilogue
    C, rows = 256, 1543
    g = torch.Generator().manual_seed(5)
    dev = "cuda:0"
    pre = torch.randn(rows, 2 * C, generator=g).to(dev, pre_dtype).requires_grad_(True)
    x1 = torch.randn(rows, C, generator=g).to(dev, x_dtype).requires_grad_(True)
    x2 = torch.randn(rows, C, generator=g).to(dev, x_dtype).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(C, generator=g)).to(dev).requires_grad_(True)
    beta = (0.2 * torch.randn(C, generator=g)).to(dev).requires_grad_(True)
    gy = torch.randn(rows, C, generator=g).to(dev, x_dtype)
    y = gate_epilogue(pre, x1, x2, gamma, beta, 1e-5)
    assert y.dtype == x_dtype
    grads = torch.autograd.grad(y, [pre, x1, x2, gamma, beta], gy)
    assert grads[0].dtype == pre_dtype and grads[1].dtype == x_dtype

    p64, a64, b64, g64, bt64 = (t.detach().double().cpu().requires_grad_(True) for t in (pre, x1, x2, gamma, beta))
    s = torch.sigmoid(p64)
    ref = torch.nn.functional.layer_norm(s[:, :C] * a64 + s[:, C:] * b64, (C,), g64, bt64, 1e-5)
    rg = torch.autograd.grad(ref, [p64, a64, b64, g64, bt64], gy.double().cpu())
    bf = 2.0 ** -8
    assert rel_err(y.detach().float().cpu().numpy(), ref.detach().numpy()) < (bf if x_dtype == torch.bfloat16 else 1e-5)
    tols = (bf if pre_dtype == torch.bfloat16 else 1e-5, bf if x_dtype == torch.bfloat16 else 1e-5,
            bf if x_dtype == torch.bfloat16 else 1e-5, 2e-5, 2e-5)
    for ours, want, tol in zip(grads, rg, tols):
        assert rel_err(ours.float().cpu().numpy(), want.numpy()) < tol


@pytest.mark.gpu
def test_gpu_gate_inference_and_autocast():
    from detrpose_b200.gate import Gate
    from detrpose_b200.functional import stats
    dev = "cuda:0"
    m = Gate(256).to(dev)
    with torch.no_grad():
        m.gate.weight.normal_(0, 0.05)
        m.gate.bias.normal_()
    x1, x2 = torch.randn(4, 9, 256, device=dev), torch.randn(4, 9, 256, device=dev)
    with torch.no_grad():
        y = m(x1, x2)
        gates = torch.sigmoid(torch.nn.functional.linear(torch.cat([x1, x2], -1), m.gate.weight, m.gate.bias))
        want = m.norm(gates[..., :256] * x1 + gates[..., 256:] * x2)
    assert rel_err(y.cpu().numpy(), want.cpu().numpy()) < 2e-5
    before = stats.get("gate_backward_launches", 0)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y2 = m(x1.requires_grad_(True), x2)
    assert y2.dtype == torch.float32                       # bf16 pre-activations, fp32 inputs -> fp32 result
    y2.sum().backward()
    assert stats["gate_backward_launches"] == before + 1
    assert rel_err(y2.detach().cpu().numpy(), want.cpu().numpy()) < 0.05
    assert x1.grad is not None and m.gate.weight.grad is not None and m.norm.weight.grad is not None


@pytest.mark.gpu
def test_gpu_gate_errors():
    from detrpose_b200 import _lib
    lib = _lib.load()
    x = torch.zeros(4, 256, device="cuda:0")
    pre = torch.zeros(4, 512, device="cuda:0")
    gam = torch.ones(256, device="cuda:0")
    y = torch.empty_like(x)
    rc = lib.msda_b200_gate_forward(pre.data_ptr(), 0, x.data_ptr(), x.data_ptr(), 0, gam.data_ptr(), gam.data_ptr(),
                                    1e-5, y.data_ptr(), None, 4, 200, None)
    assert rc == -2 or rc < 0
    assert b"C=200" in lib.msda_b200_last_error()
    rc = lib.msda_b200_gate_forward(None, 0, x.data_ptr(), x.data_ptr(), 0, gam.data_ptr(), gam.data_ptr(),
                                    1e-5, y.data_ptr(), None, 4, 256, None)
    assert rc < 0 and b"NULL" in lib.msda_b200_last_error()


# ----------------------------------------------------------------------------------------------
# patching an existing (reference-shaped) Gate class
# ----------------------------------------------------------------------------------------------
def _reference_shaped_module():
    """A module namespace holding a class with the reference Gate's structure (transformer.py:222-235),
    as `detrpose_b200.gate.install_gate` expects to find it."""
    import types

    class Gate(torch.nn.Module):
        def __init__(self, d_model):
            super().__init__()
            self.gate = torch.nn.Linear(2 * d_model, 2 * d_model)
            self.norm = torch.nn.LayerNorm(d_model)

        def forward(self, x1, x2):
            gates = torch.sigmoid(self.gate(torch.cat([x1, x2], dim=-1)))
            g1, g2 = gates.chunk(2, dim=-1)
            return self.norm(g1 * x1 + g2 * x2)

    return types.SimpleNamespace(Gate=Gate)


def test_install_gate_cpu_routes_to_original():
    from detrpose_b200.gate import install_gate, uninstall_gate
    ns = _reference_shaped_module()
    original = ns.Gate.forward
    m = ns.Gate(128)
    x1, x2 = torch.randn(3, 128), torch.randn(3, 128)
    want = m(x1, x2)
    install_gate(ns)
    install_gate(ns)                                       # idempotent
    assert ns.Gate.forward is not original
    assert torch.equal(m(x1, x2), want)                    # CPU tensors: the reference's own forward
    uninstall_gate(ns)
    assert ns.Gate.forward is original


@pytest.mark.gpu
def test_gpu_install_gate_matches_unpatched():
    from detrpose_b200.gate import install_gate, uninstall_gate
    from detrpose_b200.functional import stats
    ns = _reference_shaped_module()
    m = ns.Gate(256).to("cuda:0")
    x1 = torch.randn(2, 7, 256, device="cuda:0", requires_grad=True)
    x2 = torch.randn(2, 7, 256, device="cuda:0", requires_grad=True)
    gy = torch.randn(2, 7, 256, device="cuda:0")
    want = m(x1, x2)
    want_g = torch.autograd.grad(want, [x1, x2, *m.parameters()], gy)
    install_gate(ns)
    try:
        before = stats.get("gate_forward_launches", 0)
        got = m(x1, x2)
        assert stats["gate_forward_launches"] == before + 1
        got_g = torch.autograd.grad(got, [x1, x2, *m.parameters()], gy)
    finally:
        uninstall_gate(ns)
    assert rel_err(got.detach().cpu().numpy(), want.detach().cpu().numpy()) < 2e-5
    for a, b in zip(got_g, want_g):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 5e-5

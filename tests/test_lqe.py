"""LQE sampler: oracle vs the real reference's golden vectors (CPU) and the sm_100a
kernels vs both (GPU, through the C ABI).

Golden vectors: tests/golden/lqe_*.npz from tests/golden/make_golden_blocks.py (the reference's own
``LQE`` class, transformer.py:263-288, with the MLP input captured by a forward hook = the sampler's
statistics).  Tolerances: 1e-5 of max|ref| for statistics, output and gradients in fp32; top-k channel
indices exact; bf16 feature maps: fp32 arithmetic on the bf16-rounded map, same 1e-5.
"""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, rel_err
from oracle import lqe_numpy as ol

CASES = sorted(os.path.basename(p)[len("lqe_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN_DIR, "lqe_*.npz")))


def load_lqe_case(name):
    z = np.load(os.path.join(GOLDEN_DIR, f"lqe_{name}.npz"))
    c = {k: z[k] for k in z.files}
    c["k"], c["nb"] = int(c["topk"]), int(c["num_body_points"])
    b, l = c["poses"].shape[:2]
    c["points"] = c["poses"].reshape(b, l * c["nb"], 2)
    n_layers = len([k for k in c if k.startswith("param_reg_conf_layers_") and k.endswith("_weight")])
    c["weights"] = [c[f"param_reg_conf_layers_{i}_weight"] for i in range(n_layers)]
    c["biases"] = [c[f"param_reg_conf_layers_{i}_bias"] for i in range(n_layers)]
    return c


def test_lqe_golden_present():
    assert len(CASES) >= 3


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("dtype,tag,tol", [(np.float32, "f32", 1e-5), (np.float64, "f64", 1e-6)])
def test_oracle_lqe_statistics(name, dtype, tag, tol):
    c = load_lqe_case(name)
    cd = np.float32 if tag == "f32" else np.float64
    stat, _ = ol.lqe_statistics(c["feat"], c["points"], c["k"], dtype, coord_dtype=cd)
    b, l = c["poses"].shape[:2]
    assert rel_err(stat.reshape(b, l, -1), c[f"stat_{tag}"]) < tol
    out = ol.lqe_forward(c["scores"], c["poses"], c["feat"], c["k"], c["weights"], c["biases"], c["nb"], dtype)
    assert rel_err(out, c[f"out_{tag}"]) < (1e-5 if tag == "f32" else 2e-6)


@pytest.mark.parametrize("name", CASES)
def test_oracle_lqe_backward(name):
    """Analytic backward of the statistics, chained through the MLP by torch autograd in fp64."""
    c = load_lqe_case(name)
    b, l = c["poses"].shape[:2]
    stat, _ = ol.lqe_statistics(c["feat"], c["points"], c["k"], np.float64, coord_dtype=np.float64)
    st = torch.from_numpy(stat.reshape(b, l, -1)).requires_grad_(True)
    x = st
    for i, (w, bias) in enumerate(zip(c["weights"], c["biases"])):
        x = torch.nn.functional.linear(x, torch.from_numpy(w).double(), torch.from_numpy(bias).double())
        if i < len(c["weights"]) - 1:
            x = torch.relu(x)
    (g_stat,) = torch.autograd.grad(x, st, torch.from_numpy(c["grad_out"]).double())
    g_feat, g_pose = ol.lqe_statistics_backward(c["feat"], c["points"], c["k"], g_stat.numpy().reshape(stat.shape),
                                                np.float64, coord_dtype=np.float64)
    assert rel_err(g_feat, c["grad_feat_f64"]) < 1e-6
    assert rel_err(g_pose.reshape(c["poses"].shape), c["grad_poses_f64"]) < 1e-6


def test_reference_init_recorded():
    assert float(load_lqe_case(CASES[0])["init_last_absmax"]) == 0.0      # transformer.py:269-270


def test_lqe_module_names_and_init():
    from detrpose_b200.lqe import LQE
    m = LQE(4, 256, 2, 17)
    assert sorted(m.state_dict()) == ["reg_conf.layers.0.bias", "reg_conf.layers.0.weight",
                                      "reg_conf.layers.1.bias", "reg_conf.layers.1.weight"]
    assert m.reg_conf.layers[0].weight.shape == (256, 85) and m.reg_conf.layers[1].weight.shape == (1, 256)
    assert float(m.reg_conf.layers[-1].weight.detach().abs().max()) == 0.0
    with pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 2, 1), torch.rand(1, 2, 34), torch.zeros(1, 128, 4, 4))


# ----------------------------------------------------------------------------------------------
# GPU
# ----------------------------------------------------------------------------------------------
def _module_from_case(c, dev):
    from detrpose_b200.lqe import LQE
    m = LQE(c["k"], c["weights"][0].shape[0], len(c["weights"]), c["nb"])
    sd = {}
    for i, (w, b) in enumerate(zip(c["weights"], c["biases"])):
        sd[f"reg_conf.layers.{i}.weight"], sd[f"reg_conf.layers.{i}.bias"] = torch.from_numpy(w), torch.from_numpy(b)
    m.load_state_dict(sd)
    return m.to(dev)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("channels_last", [False, True])
def test_gpu_lqe_module_golden(name, channels_last):
    from detrpose_b200 import _lib
    from detrpose_b200 import functional as MF
    c = load_lqe_case(name)
    dev = "cuda:0"
    m = _module_from_case(c, dev)
    feat = torch.from_numpy(c["feat"]).to(dev)
    if channels_last:
        feat = feat.contiguous(memory_format=torch.channels_last)
    feat.requires_grad_(True)
    poses = torch.from_numpy(c["poses"]).to(dev).requires_grad_(True)
    scores = torch.from_numpy(c["scores"]).to(dev).requires_grad_(True)
    prev_mode = MF.get_default_coord_mode()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    MF.set_default_coord_mode(_lib.COORD_UNFUSED)          # the golden vectors come from the CPU op chain
    try:
        captured = {}
        h = m.reg_conf.register_forward_hook(lambda mod, inp, out: captured.__setitem__("stat", inp[0].detach()))
        out = m(scores, poses, feat)
        h.remove()
        grads = torch.autograd.grad(out, [feat, poses, scores, *m.parameters()], torch.from_numpy(c["grad_out"]).to(dev))
    finally:
        MF.set_default_coord_mode(prev_mode)
        torch.backends.cuda.matmul.allow_tf32 = prev
    for tag in ("f32", "f64"):
        assert rel_err(captured["stat"].cpu().numpy(), c[f"stat_{tag}"]) < 1e-5
        assert rel_err(out.detach().cpu().numpy(), c[f"out_{tag}"]) < 1e-5
        assert rel_err(grads[0].cpu().numpy(), c[f"grad_feat_{tag}"]) < 1e-5
        assert rel_err(grads[1].cpu().numpy(), c[f"grad_poses_{tag}"]) < 1e-5
        assert rel_err(grads[2].cpu().numpy(), c[f"grad_scores_{tag}"]) < 1e-5
        for (n, _), g in zip(m.named_parameters(), grads[3:]):
            assert rel_err(g.cpu().numpy(), c[f"grad_{n.replace('.', '_')}_{tag}"]) < 2e-5


@pytest.mark.gpu
@pytest.mark.parametrize("C,k", [(128, 1), (256, 4), (384, 4), (512, 8)])
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_gpu_lqe_statistics_vs_oracle_and_device_ops(C, k, dtype):
    """Model-sized map against the oracle (fp64 on the same, possibly bf16-rounded, map; indices exact) and
    against the reference's op sequence on the device (grid_sample + topk)."""
    from detrpose_b200.lqe import lqe_statistics
    from detrpose_b200 import _lib
    dev = "cuda:0"
    g = torch.Generator().manual_seed(C + k)
    B, P, hf, wf = 3, 60 * 17, 20, 24
    feat = torch.randn(B, C, hf, wf, generator=g).to(dev, dtype).requires_grad_(True)
    poses = (torch.rand(B, P, 2, generator=g) * 1.2 - 0.1).to(dev).requires_grad_(True)
    gs = torch.randn(B, P, k + 1, generator=g).to(dev)
    stat = lqe_statistics(feat, poses, k, coord_mode=_lib.COORD_FMA)

    f64 = feat.detach().float().cpu().numpy()
    want, idx = ol.lqe_statistics(f64, poses.detach().cpu().numpy(), k, np.float64, coord_dtype=np.float32)
    assert rel_err(stat.detach().cpu().numpy(), want) < 1e-5
    # Which channel a gradient goes to depends on the ORDER of the sampled values; where two of the k+1
    # largest are closer than fp32 resolution allows to tell apart, fp32 and fp64 may order them
    # differently.  Those keypoints (a handful at most) get a zero upstream gradient in this comparison.
    top_next, _ = ol.lqe_statistics(f64, poses.detach().cpu().numpy(), min(k + 1, C), np.float64, coord_dtype=np.float32)
    gaps = -np.diff(top_next[..., :-1], axis=-1)
    near_tie = (gaps < 2e-6 * np.abs(want).max()).any(-1) & (np.abs(top_next).max(-1) > 0)   # all-zero: no gradient anyway
    assert near_tie.mean() < 0.01
    gs = gs * torch.from_numpy(~near_tie).to(dev)[..., None]
    g_feat, g_pose = torch.autograd.grad(stat, [feat, poses], gs)
    wf_, wp_ = ol.lqe_statistics_backward(f64, poses.detach().cpu().numpy(), k, gs.cpu().numpy(), np.float64,
                                          coord_dtype=np.float32)
    tol_feat = 2.0 ** -8 if dtype == torch.bfloat16 else 1e-5          # grad_feat is returned in feat's dtype
    assert g_feat.dtype == dtype
    assert rel_err(g_feat.float().cpu().numpy(), wf_) < tol_feat
    # the pose gradient is discontinuous where a pixel coordinate is integral: compare away from the kinks
    diff = np.abs(g_pose.cpu().numpy() - wp_)
    assert np.mean(diff > 1e-5 * np.abs(wp_).max()) < 1e-4

    # the reference's ops on this device
    fr = feat.detach().float().requires_grad_(True)
    pr = poses.detach().clone().requires_grad_(True)
    v = torch.nn.functional.grid_sample(fr, (2 * pr - 1).view(B, P, 1, 2), mode="bilinear", padding_mode="zeros",
                                        align_corners=False).permute(0, 2, 3, 1).reshape(B, P, C)
    top = v.topk(k, dim=-1)[0]
    ref = torch.cat([top, top.mean(-1, keepdim=True)], -1)
    assert rel_err(stat.detach().cpu().numpy(), ref.detach().cpu().numpy()) < 1e-5
    rf, rp = torch.autograd.grad(ref, [fr, pr], gs)
    assert rel_err(g_feat.float().cpu().numpy(), rf.cpu().numpy()) < tol_feat
    assert np.mean(np.abs((g_pose - rp).cpu().numpy()) > 1e-5 * float(rp.abs().max())) < 1e-4


@pytest.mark.gpu
def test_gpu_lqe_all_outside_and_inference():
    from detrpose_b200.lqe import lqe_statistics
    from detrpose_b200.functional import stats
    dev = "cuda:0"
    feat = torch.randn(1, 128, 5, 5, device=dev)
    poses = torch.full((1, 7, 2), 3.0, device=dev)                    # far outside: every corner dropped
    before = stats.get("lqe_backward_launches", 0)
    with torch.no_grad():
        st = lqe_statistics(feat, poses, 4)
    assert torch.count_nonzero(st) == 0
    assert stats.get("lqe_backward_launches", 0) == before
    poses.requires_grad_(True)
    st = lqe_statistics(feat, poses, 4)
    (gp,) = torch.autograd.grad(st.sum(), poses)
    assert torch.count_nonzero(gp) == 0


@pytest.mark.gpu
def test_gpu_install_lqe_matches_unpatched():
    import types
    from detrpose_b200.lqe import install_lqe, uninstall_lqe, _MLP

    class LQE(torch.nn.Module):                    # the reference's structure (transformer.py:263-288), restated
        def __init__(self, topk, hidden_dim, num_layers, num_body_points):
            super().__init__()
            self.k, self.num_body_points = topk, num_body_points
            self.reg_conf = _MLP(num_body_points * (topk + 1), hidden_dim, 1, num_layers)

        def forward(self, scores, pred_poses, feat):
            B, L = pred_poses.shape[:2]
            pts = pred_poses.reshape(B, L, self.num_body_points, 2)
            v = torch.nn.functional.grid_sample(feat, 2 * pts - 1, mode="bilinear", padding_mode="zeros",
                                                align_corners=False).permute(0, 2, 3, 1)
            top = v.topk(self.k, dim=-1)[0]
            stat = torch.cat([top, top.mean(dim=-1, keepdim=True)], dim=-1)
            return scores + self.reg_conf(stat.reshape(B, L, -1))

    ns = types.SimpleNamespace(LQE=LQE)
    dev = "cuda:0"
    m = ns.LQE(4, 64, 2, 17).to(dev)
    feat = torch.randn(2, 256, 16, 16, device=dev, requires_grad=True)
    poses = torch.rand(2, 9, 34, device=dev, requires_grad=True)
    scores = torch.randn(2, 9, 1, device=dev)
    want = m(scores, poses, feat)
    wg = torch.autograd.grad(want.sum(), [feat, poses])
    # CPU tensors keep the original forward
    install_lqe(ns)
    try:
        got = m(scores, poses, feat)
        gg = torch.autograd.grad(got.sum(), [feat, poses])
        cpu = m.cpu()(scores.cpu(), poses.detach().cpu(), feat.detach().cpu())
    finally:
        uninstall_lqe(ns)
    assert rel_err(got.detach().cpu().numpy(), want.detach().cpu().numpy()) < 1e-5
    assert rel_err(cpu.detach().numpy(), want.detach().cpu().numpy()) < 1e-4
    assert rel_err(gg[0].cpu().numpy(), wg[0].cpu().numpy()) < 1e-5
    assert np.mean(np.abs((gg[1] - wg[1]).cpu().numpy()) > 1e-5 * float(wg[1].abs().max())) < 1e-3


@pytest.mark.gpu
def test_gpu_lqe_errors():
    from detrpose_b200 import _lib
    lib = _lib.load()
    feat = torch.zeros(1, 100, 4, 4, device="cuda:0")
    poses = torch.zeros(1, 3, 2, device="cuda:0")
    stat = torch.zeros(1, 3, 5, device="cuda:0")
    rc = lib.msda_b200_lqe_forward(feat.data_ptr(), 0, _lib.i64_array(feat.stride()), poses.data_ptr(), stat.data_ptr(),
                                   None, 1, 100, 4, 4, 3, 4, 1, None)
    assert rc < 0 and b"C=100" in lib.msda_b200_last_error()
    rc = lib.msda_b200_lqe_forward(feat.data_ptr(), 0, _lib.i64_array(feat.stride()), poses.data_ptr(), None,
                                   None, 1, 128, 4, 4, 3, 4, 1, None)
    assert rc < 0 and b"NULL" in lib.msda_b200_last_error()

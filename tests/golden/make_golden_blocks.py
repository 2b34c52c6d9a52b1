"""Golden fixtures for the blocks either side of the sampler, from the REAL
reference classes (build container only):

    python tests/golden/make_golden_blocks.py     # needs /root/reference, writes tests/golden/{gate,lqe}_*.npz

``Gate`` and ``LQE`` live in /root/reference/src/models/detrpose/transformer.py, whose package
``__init__`` files import libraries this image does not have; the module is therefore imported through
empty stand-in packages (only ``__path__`` set) and a stub ``omegaconf`` -- the classes themselves run
unmodified.  fp32 runs are the parity target, fp64 runs the high-precision arbiter (stored as fp32).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"

GATE_CASES = {
    # name: leading shape, d_model; "random" weights are stored, "structured" ones are rebuilt by the tests
    "n_like": dict(lead=(2, 5), C=128, weight="random"),
    "s_like": dict(lead=(2, 3, 6), C=256, weight="structured"),
    "x_like": dict(lead=(7,), C=384, weight="structured"),
}


def structured_weight(C: int) -> np.ndarray:
    """A (2C, 2C) gate weight made of exactly representable values by integer arithmetic, so the large
    fixtures need not store it: banded (every 5th diagonal), entries k/64 with k in -8..8."""
    i = np.arange(2 * C, dtype=np.int64)[:, None]
    j = np.arange(2 * C, dtype=np.int64)[None, :]
    k = (i * 7 + j * 13) % 17 - 8
    return (k * (np.abs(i - j) % 5 == 0)).astype(np.float32) / np.float32(64.0)


def probe_vectors(C: int):
    """Deterministic left / right probes for projecting the (2C, 2C) weight gradient of the large cases."""
    t = np.arange(2 * C, dtype=np.float64)
    return np.cos(0.37 * t + 0.1).astype(np.float32), np.sin(0.23 * t + 0.4).astype(np.float32)


def load_reference_transformer():
    def fake_pkg(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    fake_pkg("src", REF_SRC)
    fake_pkg("src.models", REF_SRC + "/models")
    fake_pkg("src.models.detrpose", REF_SRC + "/models/detrpose")
    fake_pkg("src.misc", REF_SRC + "/misc")
    oc = types.ModuleType("omegaconf")
    oc.OmegaConf = type("OmegaConf", (), {})
    oc.DictConfig = dict
    sys.modules["omegaconf"] = oc
    import src.models.detrpose.transformer as T          # noqa: E402  (the unmodified reference module)
    return T


def run_gate(T, cfg, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    C, lead = cfg["C"], cfg["lead"]
    gate = T.Gate(C)
    init = {k: v.detach().clone() for k, v in gate.state_dict().items()}
    with torch.no_grad():
        if cfg["weight"] == "random":
            gate.gate.weight.copy_(torch.randn(2 * C, 2 * C, generator=g) / (2 * C) ** 0.5)
        else:
            gate.gate.weight.copy_(torch.from_numpy(structured_weight(C)))
        gate.gate.bias.copy_(torch.randn(2 * C, generator=g))
        gate.norm.weight.copy_(1 + 0.3 * torch.randn(C, generator=g))
        gate.norm.bias.copy_(0.3 * torch.randn(C, generator=g))
    x1 = torch.randn(*lead, C, generator=g)
    x2 = 2.0 * torch.randn(*lead, C, generator=g)
    gy = torch.randn(*lead, C, generator=g)
    params = {k: v.detach().clone() for k, v in gate.state_dict().items()}
    gate = gate.to(dtype)
    a, b = x1.to(dtype).requires_grad_(True), x2.to(dtype).requires_grad_(True)
    y = gate(a, b)
    names = [n for n, _ in gate.named_parameters()]
    grads = torch.autograd.grad(y, [a, b, *gate.parameters()], gy.to(dtype))
    out = {"y": y, "grad_x1": grads[0], "grad_x2": grads[1]}
    for n, t in zip(names, grads[2:]):
        out["grad_" + n.replace(".", "_")] = t
    inputs = dict(x1=x1, x2=x2, grad_y=gy, eps=np.float32(gate.norm.eps))
    inputs.update({"param_" + k.replace(".", "_"): v for k, v in params.items()})
    inputs["init_gate_bias"] = init["gate.bias"]
    inputs["init_gate_weight_absmax"] = init["gate.weight"].abs().max()
    if cfg["weight"] != "random":                     # keep the big fixtures small: no (2C, 2C) arrays
        del inputs["param_gate_weight"]
        left, right = (torch.from_numpy(v).to(dtype) for v in probe_vectors(C))
        gw = out.pop("grad_gate_weight")
        out["grad_gate_weight_right"] = gw @ right
        out["grad_gate_weight_left"] = left @ gw
    return inputs, out


LQE_CASES = {
    # name: B, L (queries), keypoints, C, map (Hf, Wf), pose range
    "small":   dict(B=2, L=3, nb=17, C=128, hw=(6, 8), lo=-0.1, hi=1.1),
    "s_like":  dict(B=1, L=4, nb=17, C=256, hw=(10, 10), lo=0.0, hi=1.0),
    "outside": dict(B=1, L=2, nb=5, C=128, hw=(4, 5), lo=-1.5, hi=2.5),
}


def run_lqe(T, cfg, seed, dtype):
    g = torch.Generator().manual_seed(seed)
    B, L, nb, C, (hf, wf) = cfg["B"], cfg["L"], cfg["nb"], cfg["C"], cfg["hw"]
    k, hidden = 4, 32
    lqe = T.LQE(k, hidden, 2, nb)
    last_init = (lqe.reg_conf.layers[-1].weight.detach().abs().max(), lqe.reg_conf.layers[-1].bias.detach().abs().max())
    with torch.no_grad():
        for layer in lqe.reg_conf.layers:
            layer.weight.copy_(torch.randn(layer.weight.shape, generator=g) / layer.weight.shape[1] ** 0.5)
            layer.bias.copy_(0.1 * torch.randn(layer.bias.shape, generator=g))
    feat = torch.randn(B, C, hf, wf, generator=g)
    poses = torch.rand(B, L, nb * 2, generator=g) * (cfg["hi"] - cfg["lo"]) + cfg["lo"]
    scores = torch.randn(B, L, 1, generator=g)
    gout = torch.randn(B, L, 1, generator=g)
    params = {kk: v.detach().clone() for kk, v in lqe.state_dict().items()}
    lqe = lqe.to(dtype)
    f, p, sc = (t.to(dtype).requires_grad_(True) for t in (feat, poses, scores))
    captured = {}
    lqe.reg_conf.register_forward_hook(lambda m, inp, out: captured.__setitem__("stat", inp[0].detach().clone()))
    out = lqe(sc, p, f)
    grads = torch.autograd.grad(out, [f, p, sc, *lqe.parameters()], gout.to(dtype))
    res = {"out": out, "stat": captured["stat"], "grad_feat": grads[0], "grad_poses": grads[1], "grad_scores": grads[2]}
    for (n, _), t in zip(lqe.named_parameters(), grads[3:]):
        res["grad_" + n.replace(".", "_")] = t
    inputs = dict(feat=feat, poses=poses, scores=scores, grad_out=gout, topk=np.int32(k), num_body_points=np.int32(nb),
                  init_last_absmax=np.float32(max(float(last_init[0]), float(last_init[1]))))
    inputs.update({"param_" + kk.replace(".", "_"): v for kk, v in params.items()})
    return inputs, res


def main():
    if not os.path.exists(REF_SRC):
        sys.exit("reference not present: golden vectors can only be regenerated in the build container")
    torch.set_num_threads(1)
    T = load_reference_transformer()
    for i, (name, cfg) in enumerate(GATE_CASES.items()):
        inputs, r32 = run_gate(T, cfg, 500 + i, torch.float32)
        _, r64 = run_gate(T, cfg, 500 + i, torch.float64)
        blob = dict(inputs)
        for tag, res in (("f32", r32), ("f64", r64)):
            for k, t in res.items():
                blob[f"{k}_{tag}"] = t.detach().to(torch.float32)
        np.savez_compressed(os.path.join(HERE, f"gate_{name}.npz"),
                            **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in blob.items()})
        print(f"gate_{name}: y {tuple(r32['y'].shape)} |y|max {r32['y'].abs().max():.4f}")
    for i, (name, cfg) in enumerate(LQE_CASES.items()):
        inputs, r32 = run_lqe(T, cfg, 900 + i, torch.float32)
        _, r64 = run_lqe(T, cfg, 900 + i, torch.float64)
        blob = dict(inputs)
        for tag, res in (("f32", r32), ("f64", r64)):
            for k, t in res.items():
                blob[f"{k}_{tag}"] = t.detach().to(torch.float32)
        np.savez_compressed(os.path.join(HERE, f"lqe_{name}.npz"),
                            **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in blob.items()})
        print(f"lqe_{name}: stat {tuple(r32['stat'].shape)} |out|max {r32['out'].abs().max():.4f}")


if __name__ == "__main__":
    main()

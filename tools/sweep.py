"""Standalone MSDeformAttn sweep (BASELINE.json configs[4]).

4 levels (80², 40², 20², 10²), 8 heads, Dh 32, 4 points, Lq in {300, 900, 1080, 1800, 3000},
batch in {1, 8, 32, 64, 128, 256}; forward + backward of the sampling core.  For every point
of the grid it times

* the sm_100a kernels on the channel-last pyramid (bf16 value/out and fp32),
* the same kernels fed the reference's strided per-level list (repack + un-repack included),
* the reference's op sequence on this GPU -- per-level ``F.grid_sample`` + stack/flatten,
  multiply, sum, autograd backward (ms_deform_attn.py:159-193) -- "the kernel to beat",

and checks the fp32 kernels against that device run (max-abs error relative to max |ref|).
One JSON line per grid point on stdout / in --out.  Timing: CUDA events around each call,
median of the timed calls; inputs are regenerated per point (seeded).
"""
import argparse
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from detrpose_b200 import synthetic                            # noqa: E402
from detrpose_b200 import functional as MF                     # noqa: E402

PEAK_GBS = 6555.5


def _peak():
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"])
    except Exception:
        return PEAK_GBS


def reference_ops(value_list, shapes, loc, attn):
    """The reference's op sequence (ms_deform_attn.py:159-193), restated: value_list[l] is
    (N*H, Dh, H_l*W_l); returns (N, Lq, H*Dh)."""
    n, lq, h, n_levels, p, _ = loc.shape
    dh = value_list[0].shape[1]
    grids = 2 * loc - 1
    sampled = []
    for l, (hl, wl) in enumerate(shapes):
        v = value_list[l].reshape(n * h, dh, hl, wl)
        g = grids[:, :, :, l].transpose(1, 2).flatten(0, 1)            # (N*H, Lq, P, 2)
        sampled.append(F.grid_sample(v, g, mode="bilinear", padding_mode="zeros", align_corners=False))
    a = attn.transpose(1, 2).reshape(n * h, 1, lq, n_levels * p)
    out = (torch.stack(sampled, dim=-2).flatten(-2) * a).sum(-1).view(n, h * dh, lq)
    return out.transpose(1, 2)


def _time(fn, warmup, reps):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def run_point(n, lq, w, args, dev):
    shapes, H, Dh, P = w["shapes"], w["H"], w["Dh"], w["P"]
    mode = MF.get_default_coord_mode()
    rec = {"N": n, "Lq": lq, "H": H, "Dh": Dh, "levels": [list(s) for s in shapes], "P": P}
    inp = synthetic.make_inputs(n, lq, H, Dh, shapes, P, seed=n * 10007 + lq, device=dev, value_dtype=torch.float32)
    loc, attn = inp["locations"], inp["attention"]

    for name, dt in (("bf16", torch.bfloat16), ("fp32", torch.float32)):
        mem = inp["memory"].to(dt)
        pyr = MF.pack_value(mem, shapes, H)
        go = inp["grad_out"].to(dt)
        e = 2 if dt == torch.bfloat16 else 4
        b_f, b_b = synthetic.algorithmic_bytes(n, lq, H, Dh, shapes, P, e_v=e, e_o=e)
        t_f = _time(lambda: MF._forward_raw(pyr, shapes, loc, attn, dt, mode), args.warmup, args.reps)
        t_b = _time(lambda: MF._backward_raw(pyr, shapes, loc, attn, go, True, True, mode), args.warmup, args.reps)
        gbs = (b_f + b_b) / ((t_f + t_b) * 1e-3) / 1e9
        rec[name] = {"fwd_ms": round(t_f, 4), "bwd_ms": round(t_b, 4), "GBps": round(gbs, 1),
                     "frac_of_peak": round(gbs / args.peak, 4)}
        del mem, pyr, go

    # the reference's hand-over: strided per-level list made by permute/flatten/split (transformer.py:1285-1286)
    mem = inp["memory"]
    S = mem.shape[1]
    if not args.no_reference:
        def make_list(m):
            v = m.view(n, S, H, Dh).permute(0, 2, 3, 1).flatten(0, 1)
            return list(v.split([h * wd for h, wd in shapes], dim=-1))

        m_ours = mem.clone().requires_grad_(True)
        l_ours, a_ours = loc.clone().requires_grad_(True), attn.clone().requires_grad_(True)
        m_ref = mem.clone().requires_grad_(True)
        l_ref, a_ref = loc.clone().requires_grad_(True), attn.clone().requires_grad_(True)
        go = inp["grad_out"]

        def ours_step():
            for t in (m_ours, l_ours, a_ours):
                t.grad = None
            MF.clear_repack_cache()
            out = MF.ms_deform_attn_core(make_list(m_ours), shapes, l_ours, a_ours, n_heads=H)
            out.backward(go)
            return out

        def ref_step():
            for t in (m_ref, l_ref, a_ref):
                t.grad = None
            out = reference_ops(make_list(m_ref), shapes, l_ref, a_ref)
            out.backward(go)
            return out

        t_list = _time(ours_step, args.warmup, args.reps)
        ref_reps = max(3, min(args.reps, int(2000.0 / max(1.0, n * lq / 1000.0))))
        torch.cuda.reset_peak_memory_stats()
        t_ref = _time(ref_step, 2, ref_reps)
        rec["strided_list_fp32"] = {"fwd_bwd_ms": round(t_list, 4)}
        rec["reference_cuda_fp32"] = {"fwd_bwd_ms": round(t_ref, 4), "reps": ref_reps,
                                      "peak_mem_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}
        rec["speedup_vs_reference_cuda"] = {
            "same_interface_fp32": round(t_ref / t_list, 2),
            "channel_last_fp32": round(t_ref / (rec["fp32"]["fwd_ms"] + rec["fp32"]["bwd_ms"]), 2),
            "channel_last_bf16": round(t_ref / (rec["bf16"]["fwd_ms"] + rec["bf16"]["bwd_ms"]), 2)}
        o1, o2 = ours_step(), ref_step()
        # grad_locations is discontinuous where a pixel coordinate is integral: compare it away from the kinks
        rec["rel_err_vs_reference_cuda"] = {
            "out": _rel(o1, o2), "grad_value": _rel(m_ours.grad, m_ref.grad),
            "grad_attention": _rel(a_ours.grad, a_ref.grad),
            "grad_locations_median_abs": float((l_ours.grad - l_ref.grad).abs().median()),
            "grad_locations_frac_gt_1e-4": float(((l_ours.grad - l_ref.grad).abs()
                                                  > 1e-4 * l_ref.grad.abs().max()).float().mean())}
    return rec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,8,32,64,128,256")
    ap.add_argument("--queries", default="300,900,1080,1800,3000")
    ap.add_argument("--workload", default="sweep4")
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--no-reference", action="store_true")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    args.peak = _peak()
    dev = "cuda:0"
    w = synthetic.WORKLOADS[args.workload]
    fh = open(args.out, "w") if args.out else None
    for n in [int(x) for x in args.batches.split(",")]:
        for lq in [int(x) for x in args.queries.split(",")]:
            rec = run_point(n, lq, w, args, dev)
            line = json.dumps(rec)
            print(line, flush=True)
            if fh:
                fh.write(line + "\n")
                fh.flush()
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()

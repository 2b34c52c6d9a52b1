#!/usr/bin/env python
"""Backward kernel variants side by side on the bench workload (GPU box): gradients of every variant against
the default gather kernel (max error relative to max|ref|) and CUDA-event time at the bench batch.

    python tools/variant_check.py [--variants 1,10,11,12,13,14] [--batch 64] [--workload detrpose_s] [--lq N]
Prints one JSON line per variant.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from detrpose_b200 import _lib, synthetic, functional as MF   # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="1,2,3,10,11,12,13,14")
    ap.add_argument("--fwd-variants", default="")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--workload", default="detrpose_s")
    ap.add_argument("--lq", type=int, default=None)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--steps", type=int, default=50)
    args = ap.parse_args()
    dev = torch.device("cuda")
    lib = _lib.load()
    w = dict(synthetic.WORKLOADS[args.workload])
    if args.lq:
        w["Lq"] = args.lq
    vdt = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    inp = synthetic.make_inputs(args.batch, w["Lq"], w["H"], w["Dh"], w["shapes"], w["P"], seed=0, device=dev,
                                value_dtype=vdt)
    shapes = inp["shapes"]
    pyr = MF.pack_value(inp["memory"], shapes, w["H"])
    loc, att, go = inp["locations"], inp["attention"], inp["grad_out"]
    cm = MF.get_default_coord_mode()
    b_f, b_b = synthetic.algorithmic_bytes(args.batch, w["Lq"], w["H"], w["Dh"], shapes, w["P"],
                                           e_v=pyr.element_size(), e_o=pyr.element_size())

    def run_bwd():
        return MF._backward_raw(pyr, shapes, loc, att, go, True, True, cm)

    def timeit(fn):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps

    lib.msda_b200_set_variant(-1, 1)
    ref = [t.clone() for t in run_bwd()]
    for v in [int(x) for x in args.variants.split(",") if x]:
        lib.msda_b200_set_variant(-1, v)
        rec = {"bwd_variant": v}
        try:
            got = run_bwd()
            torch.cuda.synchronize()
            rec["err"] = [float((g.float() - r.float()).abs().max() / r.float().abs().max()) for g, r in zip(got, ref)]
            ms = timeit(run_bwd)
            rec["ms"] = round(ms, 4)
            rec["GBps"] = round(b_b / ms / 1e6, 1)
        except Exception as e:                      # noqa: BLE001
            rec["error"] = str(e)[:200]
        print(json.dumps(rec), flush=True)
    lib.msda_b200_set_variant(-1, -1)
    if args.fwd_variants:
        ref_out = MF._forward_raw(pyr, shapes, loc, att, vdt, cm).clone()
        for v in [int(x) for x in args.fwd_variants.split(",") if x]:
            lib.msda_b200_set_variant(v, -1)
            rec = {"fwd_variant": v}
            try:
                got = MF._forward_raw(pyr, shapes, loc, att, vdt, cm)
                torch.cuda.synchronize()
                rec["err"] = float((got.float() - ref_out.float()).abs().max() / ref_out.float().abs().max())
                ms = timeit(lambda: MF._forward_raw(pyr, shapes, loc, att, vdt, cm))
                rec["ms"] = round(ms, 4)
                rec["GBps"] = round(b_f / ms / 1e6, 1)
            except Exception as e:                  # noqa: BLE001
                rec["error"] = str(e)[:200]
            print(json.dumps(rec), flush=True)
        lib.msda_b200_set_variant(-1, -1)


if __name__ == "__main__":
    main()

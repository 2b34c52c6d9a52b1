#!/usr/bin/env python
"""Layout hand-over kernels (row f2) against the HBM roofline (GPU box): repack of the reference's value list
(N*H, Dh, H_l*W_l) -> channel-last pyramid, and the inverse for the fp32 gradient.  One JSON line per dtype."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from detrpose_b200 import functional as MF, synthetic                 # noqa: E402
from oracle import msda_torch as otorch                               # noqa: E402  (builds the reference's list)


def timeit(fn, steps=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def main():
    peak = 6555.5
    try:
        peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                           "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:                                                   # noqa: BLE001
        pass
    w = synthetic.WORKLOADS["detrpose_s"]
    for N, vdt in ((64, torch.bfloat16), (64, torch.float32), (16, torch.bfloat16)):
        inp = synthetic.make_inputs(N, w["Lq"], w["H"], w["Dh"], w["shapes"], w["P"], seed=0, device="cuda", value_dtype=vdt)
        value = otorch.make_value_list(inp["memory"], w["H"], w["shapes"])        # the reference's strided copy
        S = inp["memory"].shape[1]
        es = inp["memory"].element_size()
        nbytes = inp["memory"].numel() * es
        MF.clear_repack_cache()

        def repack():
            MF.clear_repack_cache()
            return MF.pack_value(value, inp["shapes"], w["H"])
        pyr = repack()
        gv = torch.randn(N, S, w["H"], w["Dh"], device="cuda")
        t_r = timeit(repack)
        t_u = timeit(lambda: MF._unpack_grad(gv, inp["shapes"], w["H"], vdt))
        t_c = timeit(lambda: inp["memory"].clone())
        rec = {"batch": N, "dtype": str(vdt).split(".")[-1],
               "repack_us": round(t_r * 1e3, 1), "repack_GBps": round(2 * nbytes / t_r / 1e6, 1),
               "unpack_grad_us": round(t_u * 1e3, 1), "unpack_grad_GBps": round((gv.numel() * 4 + nbytes) / t_u / 1e6, 1),
               "torch_clone_same_bytes_us": round(t_c * 1e3, 1), "hbm_peak_GBps": peak}
        print(json.dumps(rec), flush=True)


if __name__ == "__main__":
    main()

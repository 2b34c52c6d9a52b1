#!/usr/bin/env python
"""Forward kernel forms side by side (GPU box): the lean kernel without its two defaults (variant 107: one lane
group per item for every row size, no L2 prefetch), the default choice without the L2 prefetch (150), the default,
and the default on a head-major copy of the pyramid ((N, H, S, Dh) storage behind the same (N, S, H, Dh) view).
Prints one JSON line per shape: time per launch and the error of each form against variant 107.
"""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from detrpose_b200 import _lib, synthetic, functional as MF
dev = torch.device("cuda")
lib = _lib.load()
def timeit(fn, steps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps
cases = (("detrpose_s", 64, torch.bfloat16, None), ("detrpose_s", 64, torch.float32, None), ("detrpose_n", 64, torch.bfloat16, None),
         ("detrpose_n", 64, torch.float32, None), ("detrpose_x", 32, torch.bfloat16, None), ("detrpose_x", 32, torch.float32, None),
         ("detrpose_s", 16, torch.bfloat16, 1584), ("detrpose_s", 1, torch.bfloat16, None), ("detrpose_n", 1, torch.bfloat16, None))
for wl, N, dt, lq in cases:
    w = dict(synthetic.WORKLOADS[wl])
    if lq: w["Lq"] = lq
    inp = synthetic.make_inputs(N, w["Lq"], w["H"], w["Dh"], w["shapes"], w["P"], seed=0, device=dev, value_dtype=dt)
    shapes = inp["shapes"]
    pyr = MF.pack_value(inp["memory"], shapes, w["H"])            # (N,S,H,Dh)
    hm = pyr.permute(0, 2, 1, 3).contiguous().permute(0, 2, 1, 3)  # same logical shape, head-major storage
    loc, att, go = inp["locations"], inp["attention"], inp["grad_out"]
    cm = MF.get_default_coord_mode()
    lib.msda_b200_set_variant(107, -1)
    ref = MF._forward_raw(pyr, shapes, loc, att, torch.float32, cm)
    rec = {"workload": wl, "batch": N, "Lq": w["Lq"], "dtype": str(dt)}
    for name, v, t in (("plain_pm", 107, pyr), ("default_nopf_pm", 150, pyr), ("default_pm", -1, pyr), ("default_hm", -1, hm)):
        lib.msda_b200_set_variant(v, -1)
        o = MF._forward_raw(t, shapes, loc, att, torch.float32, cm)
        rec[name + "_err"] = float((o - ref).abs().max() / ref.abs().max())
        rec[name + "_ms"] = round(timeit(lambda: MF._forward_raw(t, shapes, loc, att, dt, cm)), 4)
    lib.msda_b200_set_variant(-1, -1)
    print(json.dumps(rec), flush=True)

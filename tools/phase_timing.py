"""Per-phase wall time inside the gather-form backward CTAs (clock64 stamps), by pyramid level."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from detrpose_b200 import synthetic, _lib                     # noqa: E402
from detrpose_b200 import functional as MF                    # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    dtype = torch.bfloat16 if (len(sys.argv) < 3 or sys.argv[2] == "bf16") else torch.float32
    dev = "cuda:0"
    w = synthetic.WORKLOADS["detrpose_s"]
    inp = synthetic.make_inputs(N, w["Lq"], w["H"], w["Dh"], w["shapes"], w["P"], seed=0, device=dev, value_dtype=dtype)
    pyr = MF.pack_value(inp["memory"], w["shapes"], w["H"])
    lib = _lib.load()
    L = len(w["shapes"])
    nctas = N * w["H"] * L
    buf = torch.zeros((nctas, 8), dtype=torch.int64, device=dev)
    need_value = not (len(sys.argv) > 3 and sys.argv[3] == "novalue")
    args = (pyr, inp["shapes"], inp["locations"], inp["attention"], inp["grad_out"], need_value, True, MF.get_default_coord_mode())
    for _ in range(3):
        MF._backward_raw(*args)
    _lib.check(lib.msda_b200_debug_phase_buffer(buf.data_ptr()), "phase buffer")
    MF._backward_raw(*args)
    torch.cuda.synchronize()
    _lib.check(lib.msda_b200_debug_phase_buffer(None), "phase buffer")
    t = buf.cpu().double()
    d = t[:, 1:7] - t[:, 0:6]
    names = ["P0 stage g", "P1 count", "P2 scan", "P3 scatter", "P4 pixels", "P5 samples"]
    out = {}
    level = buf[:, 7].cpu()                       # the kernel records each CTA's level in slot 7
    for l in range(L):
        m = level == l
        out[f"level{l}"] = {n: round(float(d[m][:, i].mean()), 0) for i, n in enumerate(names)}
        out[f"level{l}"]["total"] = round(float((t[m][:, 6] - t[m][:, 0]).mean()), 0)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

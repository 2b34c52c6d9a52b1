#!/usr/bin/env python
"""SASS evidence for profiles/: opcode histograms and the first FFMA2 region of the two DEFAULT kernels in
detrpose_b200/libmsda_b200.so (cuobjdump -sass), i.e. what `msda_b200_forward` / `msda_b200_backward` launch for the
bench configuration (bf16 value and output, Dh 32).

    python tools/sass_excerpt.py > profiles/r02_sass_default_kernels.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "detrpose_b200", "libmsda_b200.so")
KERNELS = [
    ("fwd_lean_kernel<G=4, K=1, bf16 value, bf16 out, one lane group per item>  (forward, default)",
     r"fwd_lean_kernelILi4ELi1ELb1ELb1ELb0EE"),
    ("bwd_gather_kernel<G=4, K=1, bf16, 1024 threads, single chunk, overwrite, unfused>  (backward, default)",
     r"bwd_gather_kernelILi4ELi1ELb1ELi1024ELb1ELb1ELb0EE"),
]


def main():
    names = subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout
    syms = sorted(set(re.findall(r"_ZN4msda\w+", names)))
    print("# SASS of the two DEFAULT kernels in detrpose_b200/libmsda_b200.so (cuobjdump -sass -fun <mangled name>), round 2.")
    print("# Opcode histogram (static instruction counts) and the hot loops.  Neither default kernel contains tensor-core")
    print("# (UTC*MMA / HMMA) or TMA tile loads (UTMALDG): the sampler is gather/scatter work at ~3 flop/byte.  The forward")
    print("# issues one bulk L2 prefetch per CTA (UBLKPF.L2 = cp.async.bulk.prefetch.L2); the opt-in staged forward")
    print("# (fwd_staged_kernel, variants 2/3) is the kernel with UTMALDG.4D.")
    for title, pat in KERNELS:
        cand = [s for s in syms if re.search(pat, s)]
        if not cand:
            print(f"\n## {title}: symbol not found ({pat})")
            continue
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", cand[0], LIB], capture_output=True, text=True).stdout
        lines = [l for l in sass.splitlines() if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l)]
        ops, forms = collections.Counter(), collections.Counter()
        for l in lines:
            body = re.sub(r"/\*.*?\*/", "", l).strip().rstrip(";").strip()
            toks = body.split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            ops[op.split(".")[0]] += 1
            if re.match(r"(LDG|STG|LDS|STS|LDGSTS|FFMA2|BAR|RED|ATOMS|SHFL|UTMALDG|UBLKPF|HMMA)", op):
                forms[op] += 1
        print(f"\n## {title}: {len(lines)} instructions\n   {cand[0]}")
        print("opcode classes: " + ", ".join(f"{k} {v}" for k, v in ops.most_common(28)))
        print("memory / packed-math forms: " + ", ".join(f"{k} {v}" for k, v in forms.most_common(20)))
        first = next((i for i, l in enumerate(lines) if "FFMA2" in l), None)
        if first is not None:
            a, b = max(first - 12, 0), min(first + 44, len(lines))
            print(f"\n### first FFMA2 region (instructions {a}..{b})")
            for l in lines[a:b]:
                print(re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l.rstrip()))


if __name__ == "__main__":
    sys.exit(main())

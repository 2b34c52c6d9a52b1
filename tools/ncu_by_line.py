#!/usr/bin/env python
"""Executed warp instructions and stall samples per CUDA source line for one kernel of an ncu report.

ncu's CSV export of the source page is per SASS instruction only; this joins it, instruction by instruction, with
the line table of the same kernel in the built library (`nvdisasm -g` on the cubin extracted by `cuobjdump -xelf`;
the library is compiled with -lineinfo) and sums per line of the kernel's own file (inlined helpers are listed
under the helper's file and line).

    python tools/ncu_by_line.py gpurun_out/r02_full.ncu-rep bwd_gather [top] > profiles/r02_bwd_by_source_line.txt
    python tools/ncu_by_line.py gpurun_out/r02_full.ncu-rep bwd_gather 12 --shared   # bank-conflict wavefronts per line
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "detrpose_b200", "libmsda_b200.so")
HEADER = []


def sass_counts(rep, pattern):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    heads = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
    for n, i in enumerate(heads):
        if re.search(pattern, rows[i][1]):
            end = heads[n + 1] if n + 1 < len(heads) else len(rows)
            global HEADER
            HEADER = rows[i + 1]
            return rows[i][1], rows[i + 2:end]
    raise SystemExit(f"no kernel matching {pattern!r} in {rep}")


def mangled_candidates(demangled):
    syms = re.findall(r"_ZN4msda\w+", subprocess.run(["cuobjdump", "-elf", LIB], capture_output=True, text=True).stdout)
    syms = sorted({s for s in syms if "_param_" not in s})
    dem = subprocess.run(["c++filt"] + syms, capture_output=True, text=True).stdout.splitlines()
    norm = lambda t: re.sub(r"\(int\)|\(bool\)|\s", "", t).replace("true", "1").replace("false", "0")   # noqa: E731
    want = norm(demangled.split("(msda::Problem")[0])
    return [s for s, d in zip(syms, dem) if norm(d.split("(msda::Problem")[0]) == want]


def line_table(symbol):
    with tempfile.TemporaryDirectory() as tmp:
        subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, capture_output=True)
        for f in sorted(os.listdir(tmp)):
            dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, f)], capture_output=True, text=True).stdout
            if f".text.{symbol}:" not in dis:
                continue
            lines = dis.split("\n")
            a = lines.index(f".text.{symbol}:")
            out, cur = [], None
            for l in lines[a + 1:]:
                if l.startswith(".text.") or l.startswith(".section"):
                    break
                m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                elif re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
                    out.append(cur)
            return out
    raise SystemExit(f"{symbol} not found in any cubin of {LIB}")


def shared_conflicts(name, table, sass, header, top):
    """Shared-memory wavefronts beyond the ideal count (bank-conflict replays), per source line and opcode."""
    iw, ie = header.index("L1 Wavefronts Shared"), header.index("L1 Wavefronts Shared Excessive")
    wav, exc, ops = collections.Counter(), collections.Counter(), collections.defaultdict(set)
    for loc, r in zip(table, sass):
        w = int(r[iw] or 0)
        if w:
            wav[loc] += w
            exc[loc] += int(r[ie] or 0)
            t = r[1].split()
            ops[loc].add(t[1] if t[0].startswith("@") else t[0])
    tw, te = sum(wav.values()), max(sum(exc.values()), 1)
    print(f"# {name.split('(msda::Problem')[0]}\n# {tw / 1e6:.1f} M shared-memory wavefronts, {te / 1e6:.1f} M of them "
          f"beyond the ideal count ({100 * te / tw:.1f} %); per source line: M excess, M total, % of all excess")
    for loc, e in exc.most_common(top):
        path = os.path.join(ROOT, "detrpose_b200", "csrc", loc[0]) if loc else ""
        text = open(path).read().split("\n")[loc[1] - 1].strip()[:90] if path and os.path.exists(path) else ""
        print(f"{e / 1e6:7.2f} {wav[loc] / 1e6:7.2f} {100 * e / te:5.1f}%  {loc[0]}:{loc[1]:<5d} {','.join(sorted(ops[loc])):22s} {text}")


def main():
    rep, pattern = sys.argv[1], sys.argv[2]
    args = [a for a in sys.argv[3:] if not a.startswith("--")]
    top = int(args[0]) if args else 60
    name, sass = sass_counts(rep, pattern)
    cands = mangled_candidates(name)
    if not cands:
        raise SystemExit(f"cannot find the symbol of {name}")
    table = line_table(cands[0])
    if len(table) != len(sass):
        raise SystemExit(f"{len(sass)} instructions in the report, {len(table)} in the library: rebuild / re-profile")
    if "--shared" in sys.argv:
        return shared_conflicts(name, table, sass, HEADER, top)
    inst, smp = collections.Counter(), collections.Counter()
    for loc, r in zip(table, sass):
        inst[loc] += int(r[5])
        smp[loc] += int(r[2])
    ti, ts = sum(inst.values()), max(sum(smp.values()), 1)
    src = {}
    print(f"# {name.split('(msda::Problem')[0]}\n# {ti / 1e6:.1f} M warp instructions, {ts} stall samples; "
          f"per source line: M instructions, % of instructions, % of samples")
    for loc, v in inst.most_common(top):
        text = ""
        if loc:
            path = os.path.join(ROOT, "detrpose_b200", "csrc", loc[0])
            if os.path.exists(path):
                src.setdefault(path, open(path).read().split("\n"))
                text = src[path][loc[1] - 1].strip()[:100]
        where = f"{loc[0]}:{loc[1]}" if loc else "?"
        print(f"{v / 1e6:8.2f} {100 * v / ti:5.1f}% {100 * smp[loc] / ts:5.1f}%  {where:28s} {text}")


if __name__ == "__main__":
    main()

"""Copy the round's measurement artefacts from gpurun_out/ into profiles/ (tracked) and condense the
ncu reports (needs `ncu` on PATH; reads gpurun_out/<tag>_full.ncu-rep and gpurun_out/<tag>_launches.csv).

    python tools/refresh_profiles.py [r02]
"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
COPY = {"bench_r01.json": "r01_bench_bf16.json", "bench_fp32_r01.json": "r01_bench_fp32.json",
        "bench_ref_r01.json": "r01_bench_reference_cpu.json", "bench_x_r01.json": "r01_bench_detrpose_x_b32.json",
        "bench_n_r01.json": "r01_bench_detrpose_n_b64.json", "bench_degen_r01.json": "r01_bench_degenerate.json",
        "bench_ltrain_r01.json": "r01_bench_detrpose_l_train_lq1584_b16.json",
        "phases_bf16_n64.json": "r01_bwd_phase_cycles_n64.json", "r01_launches.csv": "r01_ncu_launch_list.csv"}
KEEP = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sectors_srcunit_tex_op_write.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    if tag == "r01":
        for src, dst in COPY.items():
            if os.path.exists(os.path.join(G, src)):
                shutil.copy(os.path.join(G, src), os.path.join(P, dst))
    elif os.path.exists(os.path.join(G, f"{tag}_launches.csv")):
        shutil.copy(os.path.join(G, f"{tag}_launches.csv"), os.path.join(P, f"{tag}_ncu_launch_list.csv"))
    rep = os.path.join(G, f"{tag}_full.ncu-rep")
    if os.path.exists(rep):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        keep = KEEP + [k for k in hdr if 'issue_stalled' in k and 'per_issue_active' in k]
        out = [["metric", "unit"] + [d[hdr.index('Kernel Name')][:60] for d in data]]
        for k in keep:
            if k in hdr:
                i = hdr.index(k)
                out.append([k, units[i]] + [d[i] for d in data])
        with open(os.path.join(P, f"{tag}_ncu_full_summary.csv"), "w", newline="") as f:
            csv.writer(f).writerows(out)

        def val(d, k):
            i = hdr.index(k)
            return float(d[i].replace(',', '')) * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(units[i], 1)
        traffic = {}
        for d in data:
            name = d[hdr.index('Kernel Name')]
            key = 'backward_dram_bytes_per_launch' if 'bwd' in name else 'forward_dram_bytes_per_launch'
            traffic[key] = int(val(d, 'dram__bytes_read.sum') + val(d, 'dram__bytes_write.sum'))
            print(name[:45], d[hdr.index('gpu__time_duration.sum')], "us", round(traffic[key] / 1e6, 1), "MB dram, issue",
                  d[hdr.index('smsp__issue_active.avg.pct_of_peak_sustained_active')], "%, inst", d[hdr.index('smsp__inst_executed.sum')])
        traffic['source'] = (f'profiles/{tag}_ncu_full_summary.csv (ncu --set full, bench.py --steps 3 --warmup 3, '
                             'batch 64, bf16)')
        json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)
    ll = os.path.join(P, f"{tag}_ncu_launch_list.csv")
    if os.path.exists(ll):
        rows = list(csv.DictReader(l for l in open(ll) if l.startswith('"')))
        out = []
        for r in rows:
            n = r['Kernel Name']
            short = n.split('(')[0][:70] if ('lean' in n or 'gather' in n) else n.split('<')[0].replace('void ', '')[-60:]
            out.append((r['ID'], short, r['Grid Size'], r['Block Size'], float(r['Metric Value']) / 1e3))
        fw = [o[4] for o in out if 'lean' in o[1]]
        bw = [o[4] for o in out if 'gather' in o[1]]
        with open(os.path.join(P, f"{tag}_ncu_launch_list_summary.txt"), "w") as f:
            f.write("# ncu --metrics gpu__time_duration.sum --clock-control none; python bench.py --steps 3 --warmup 3 "
                    "--no-e2e --no-cpu (batch 64, bf16)\n# cold-cache serialised per-launch times: compare SHARES. "
                    "id, kernel, grid, block, us\n")
            for o in out:
                f.write(f"{o[0]:>3s}  {o[1]:72s} {o[2]:>14s} {o[3]:>14s} {o[4]:10.1f}\n")
            if fw and bw:
                tf, tb = sum(fw) / len(fw), sum(bw) / len(bw)
                bench = json.load(open(os.path.join(P, f"{tag}_bench_bf16.json")))["kernels"]
                share = bench["backward_ms"] / (bench["backward_ms"] + bench["forward_ms"]) * 100
                f.write(f"# mean forward {tf:.1f} us, mean backward {tb:.1f} us -> backward share of a step "
                        f"{tb / (tf + tb) * 100:.1f}% (bench.py CUDA events: {share:.1f}%)\n")
        print(open(os.path.join(P, f"{tag}_ncu_launch_list_summary.txt")).read().splitlines()[-1])


if __name__ == "__main__":
    main()

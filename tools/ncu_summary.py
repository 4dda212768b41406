"""Summarise an .ncu-rep (read here, no GPU): key metrics per launch + instruction mix + SIMT efficiency.
usage: python tools/ncu_summary.py file.ncu-rep [--json out.json]"""
import collections
import csv
import io
import json
import subprocess
import sys

rep = sys.argv[1]
KEYS = ['gpu__time_duration.sum', 'SM_B.TriageCompute.l1tex__t_sectors.sum', 'SM_B.TriageCompute.l1tex__t_sectors_lookup_hit.sum',
        'SM_B.TriageCompute.l1tex__t_sectors_lookup_miss.sum', 'derived__l1tex__lsu_writeback_bytes_mem_lgds.sum.per_second',
        'lts__t_sectors.sum', 'lts__t_sectors_op_read.sum', 'lts__t_sectors_srcunit_tex.sum', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__inst_executed.sum', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_active',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active']
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
out = []
for r in rows[2:]:
    d = {"kernel": r[hdr.index("Kernel Name")]}
    for k in KEYS:
        if k in hdr:
            d[k] = (r[hdr.index(k)] + " " + units[hdr.index(k)]).strip()
    out.append(d)
    print("==", d["kernel"])
    for k in KEYS:
        if k in d:
            print(f"  {k:86s} {d[k]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
blk = rows[starts[0] + 1: starts[1] if len(starts) > 1 else None]
H, data = blk[0], blk[1:]
ci, ti, si, sa = H.index("Instructions Executed"), H.index("Thread Instructions Executed"), H.index("Source"), H.index("# Samples")
tot_i = sum(int(r[ci]) for r in data)
tot_t = sum(int(r[ti]) for r in data)
op, opt = collections.Counter(), collections.Counter()
for r in data:
    s = r[si].strip().split()
    o = (s[0] if not s[0].startswith('@') else s[1]).split('.')[0]
    op[o] += int(r[ci])
    opt[o] += int(r[ti])
print(f"warp instructions {tot_i}, threads/instruction {tot_t / tot_i:.2f}, SASS lines {len(data)}")
mix = {}
for k, v in op.most_common(16):
    mix[k] = {"warp_inst": v, "pct": round(100 * v / tot_i, 1), "avg_threads": round(opt[k] / max(v, 1), 1)}
    print(f"  {k:10s} {v:12d} {100 * v / tot_i:5.1f}%  avg threads {opt[k] / max(v, 1):.1f}")
stall_cols = [c for c in H if c.startswith("stall_") and "Not Issued" not in c]
st = {c: sum(int(r[H.index(c)] or 0) for r in data) for c in stall_cols}
tot_s = sum(st.values())
print("stall samples:", ", ".join(f"{k[6:]} {100 * v / tot_s:.1f}%" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
if "--json" in sys.argv:
    json.dump({"report": rep, "launches": out, "warp_instructions": tot_i, "threads_per_instruction": tot_t / tot_i,
               "instruction_mix": mix, "stall_samples_pct": {k[6:]: round(100 * v / tot_s, 1) for k, v in st.items() if v}},
              open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
if "--hot" in sys.argv:
    for idx, r in enumerate(data):
        n = int(r[ci])
        if n > tot_i * 0.004:
            print(idx, r[si].strip()[:64].ljust(64), n, f"{int(r[ti]) / n:.1f}", r[sa])

#!/bin/bash
# After `gpurun -- bash tests/scripts/gpu_r2_round.sh`: copy the round-2 evidence from gpurun_out/ into profiles/
# (bench lines, launch list of the bench command, ncu --set full summary of the segment kernel on all of C5,
# traffic.json = what bench.py reads into roofline.traffic / roofline.on_chip).
set -e
cd "$(dirname "$0")/../.."
cp gpurun_out/r2_bench_n1.json profiles/r2_bench_n1.json
cp gpurun_out/r2_bench_ref.json profiles/r2_bench_reference_arm.json
[ -f gpurun_out/r2_bench_n2.json ] && cp gpurun_out/r2_bench_n2.json profiles/r2_bench_n2.json
[ -f gpurun_out/r2_bench_n4.json ] && cp gpurun_out/r2_bench_n4.json profiles/r2_bench_n4.json
[ -f gpurun_out/r2_bench_n8.json ] && cp gpurun_out/r2_bench_n8.json profiles/r2_bench_n8.json
{ echo "ncu --metrics gpu__time_duration.sum --clock-control none of: python bench.py --steps 2 --warmup 3 --no-cpu-baseline (B200, N = 1)";
  echo "per-kernel totals (times under ncu are serialised and cold-cache: shares, not absolutes)"; echo;
  python tests/ncu_summary.py --launches gpurun_out/r2_launches.csv; } > profiles/r2_seg_bench_launches.txt
ncu -i gpurun_out/r2_seg_c5.ncu-rep --page raw --csv > /tmp/raw_c5.csv 2>/dev/null
{ echo "ncu --set full --clock-control none --import-source on -k regex:x3_seg -s 1 -c 1: x3_seg_kernel on ALL of C5 (211 938 580 positions,";
  echo "-t 15 -w 8), device resident, one launch = the launch bench.py's device leg times (tests/gpu_c5_device.py).  B200.";
  echo; python tests/ncu_summary.py /tmp/raw_c5.csv; } > profiles/r2_seg_ncu_summary.txt
ncu -i gpurun_out/r2_seg_c5.ncu-rep --page source --csv --print-source cuda,sass > /tmp/src_c5.csv 2>/dev/null
{ echo "hottest source lines of x3_seg_kernel on all of C5 (same capture): stall samples, executed warp instructions,";
  echo "shared-memory wavefronts (+ excessive = bank conflicts), top two stall reasons"; echo;
  python tests/ncu_lines.py /tmp/src_c5.csv x3_search_seg.cu 60; } > profiles/r2_seg_hot_lines.txt
python - <<'PY'
import csv, json
rows = list(csv.reader(open('/tmp/raw_c5.csv')))
hdr, vals = rows[0], rows[2]
m = dict(zip(hdr, vals))
units = dict(zip(hdr, rows[1]))
def f(k):
    v = float(m[k].replace(',', ''))
    u = units.get(k, '')
    return v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}.get(u, 1)
j = {"kernel": "x3_seg_kernel",
     "dram_bytes_per_launch": int(f("dram__bytes_read.sum") + f("dram__bytes_write.sum")),
     "source": "profiles/r2_seg_ncu_summary.txt: dram__bytes_read.sum + dram__bytes_write.sum of the one x3_seg_kernel launch over all "
               "of C5 (211 938 580 positions), ncu --set full",
     "on_chip": {"issue_slots_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 "shared_memory_wavefronts_pct_of_peak": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
                 "shared_memory_bank_conflict_share": f("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum") / f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                 "warps_eligible_per_cycle": f("smsp__warps_eligible.avg.per_cycle_active"),
                 "stalled_barrier_per_issue": f("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"),
                 "stalled_short_scoreboard_per_issue": f("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
                 "warp_instructions": f("smsp__inst_executed.sum"),
                 "duration_us_under_ncu": f("gpu__time_duration.sum") / 1e3 if units.get("gpu__time_duration.sum") == "ns" else f("gpu__time_duration.sum") * (1e3 if units.get("gpu__time_duration.sum") == "ms" else 1),
                 "measured_instruction_mix_peaks": "profiles/r2_ubench.json",
                 "source": "same capture"}}
json.dump(j, open('profiles/traffic.json', 'w'), indent=1)
d = json.load(open('profiles/r2_bench_n1.json'))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "plugin", d["e2e"]["plugin"]["ms_per_call"],
      "roofline frac", d["roofline"]["frac"], "traffic", j["dram_bytes_per_launch"], "vs algorithmic", d["roofline"]["algorithmic_bytes_per_launch"])
PY

#!/bin/bash
# After `gpurun -- bash tests/scripts/gpu_round.sh`: copy the evidence from gpurun_out/ into profiles/.
set -e
cd "$(dirname "$0")/../.."
cp gpurun_out/bench_n1.json profiles/r1_bench_n1_rank.json
cp gpurun_out/bench_ref.json profiles/r1_bench_reference_arm.json
{ head -1 profiles/r1_rank_bench_launches.txt; python tests/ncu_summary.py --launches gpurun_out/launches.csv; } > /tmp/launches.txt
cp /tmp/launches.txt profiles/r1_rank_bench_launches.txt
ncu -i gpurun_out/rank_full.ncu-rep --page raw --csv > /tmp/raw.csv 2>/dev/null
{ head -3 profiles/r1_rank_ncu_summary.txt; echo "(final tree of the round)"; echo; python tests/ncu_summary.py /tmp/raw.csv; } > /tmp/ncu_sum.txt
cp /tmp/ncu_sum.txt profiles/r1_rank_ncu_summary.txt
python - <<'PY'
import re, json
t = open('profiles/r1_rank_ncu_summary.txt').read()
vals = []
for b in t.split("\n== ")[1:]:
    if "radix" in b.split("\n")[0]:
        r = float(re.search(r"dram__bytes_read.sum \[Mbyte\] = ([\d.]+)", b).group(1))
        w = float(re.search(r"dram__bytes_write.sum \[Mbyte\] = ([\d.]+)", b).group(1))
        vals.append((r + w) * 1e6)
j = json.load(open('profiles/traffic.json'))
j["dram_bytes_per_launch"] = int(sum(vals) / len(vals))
json.dump(j, open('profiles/traffic.json', 'w'))
d = json.load(open('profiles/r1_bench_n1_rank.json'))
t = open('profiles/r1_rank_bench_launches.txt').read()
tot = rad = 0.0
for m in re.finditer(r"x\s+([\d.]+) ms\s+[\d.]+ %.*?(x3_rank_\w+)", t):
    tot += float(m.group(1))
    rad += float(m.group(1)) if "radix" in m.group(2) else 0.0
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline frac", d["roofline"]["frac"], "share events",
      d["roofline"]["share_of_search"], "share ncu", rad / tot, "compress", d["compress"]["value"],
      d["compress"]["value_excluding_cuda_startup"], d["compress"]["cuda_startup_s"])
PY

#!/bin/bash
# usage: scripts/ncu_launches.sh <tag> [env assignments...]  -- per-launch device time + SM clock
tag=$1; shift
env "$@" ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_$tag.log 2>&1
python - "$tag" <<'PY'
import csv, collections, re, sys
tag = sys.argv[1]
lines = [l for l in open(f"gpurun_out/launches_{tag}.csv") if not l.startswith("==")]
agg = collections.OrderedDict()
for row in csv.DictReader(lines):
    name = re.sub(r"\(.*", "", row["Kernel Name"])[:60]
    v = float(row["Metric Value"].replace(",", ""))
    a = agg.setdefault(name, {})
    a.setdefault(row["Metric Name"], []).append(v)
print("==", tag)
for k, m in agg.items():
    t = m.get("gpu__time_duration.sum", [0]); c = m.get("sm__cycles_elapsed.avg.per_second", [0])
    print(f"{len(t):4d} x {sum(t)/len(t)/1e3:10.1f} us  clk {sum(c)/len(c):6.3f}  {k}")
PY

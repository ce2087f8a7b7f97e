#!/bin/bash
# session 21: refresh of the committed cbox lines / ncu summaries on the final build
out=gpurun_out; mkdir -p $out; tag=${1:-s21}
timeout 900 python bench.py > $out/${tag}_bench_default.json 2> $out/${tag}_bench_default.err; cut -c1-200 $out/${tag}_bench_default.json
timeout 900 python bench.py --workload cbox --bvh-counters --cpu-baseline 0 > $out/${tag}_bench_cbox.json 2> $out/${tag}_bench_cbox.err; cut -c1-120 $out/${tag}_bench_cbox.json
NCU="ncu --set full --clock-control none --import-source on"
prof() { local name=$1 k=$2 skip=$3; shift 3
  timeout 600 $NCU -k regex:$k -s $skip -c 1 -o /tmp/${name} -f "$@" > $out/${tag}_ncu_${name}.log 2>&1
  ncu -i /tmp/${name}.ncu-rep --page raw --csv > $out/${tag}_ncu_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${name}.ncu-rep --page source --csv > $out/${tag}_ncu_${name}_src.csv 2>/dev/null; rm -f /tmp/${name}.ncu-rep; }
prof cbox_trace 'k_trace' 60 python scripts/render_once.py --workload cbox --spp 16
prof cbox_shade 'k_shade' 60 python scripts/render_once.py --workload cbox --spp 16
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $out/${tag}_launches_cbox.csv python scripts/render_once.py --workload cbox --spp 32 --flags 0 > $out/${tag}_ncu_launches.log 2>&1
du -sh $out

#!/bin/bash
# session 19 (final build of round 2, second half): the whole GPU suite (incl. parity at size), the driver's default bench line,
# the four workload lines, the reference arm, ncu -> CSV of the stage kernels, launch list
out=gpurun_out; mkdir -p $out; tag=${1:-s19}
timeout 1800 python -m pytest tests -m gpu -q -s > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
grep -E "^PARITY|LBVH|passed|failed|exit|^E  |Error|^FAILED" $out/${tag}_pytest.log | cut -c1-420 | tail -24
show() {
  python - $1 $2 <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = j["roofline"]
    print(f"{sys.argv[2]:22s} {j['value']:8.1f} Msamples/s e2e {j['e2e']['value']:8.1f} step {j['ms_per_step']:7.2f} ms  " +
          "  ".join(f"{k} {v['ms']:.1f} ms {v['frac']:.3f}" for k, v in r["stages"].items()) + f"  step_frac {r['step_frac']:.3f} clk {j['clocks']['sm_mhz']}", j.get("bvh", {}).get("nodes_per_ray"), j.get("bvh", {}).get("tris_per_ray"))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run() { local name=$1; shift 1; timeout 900 python bench.py "$@" > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err; show $out/${tag}_bench_$name.json $name; }
run default
run cbox --workload cbox --bvh-counters --cpu-baseline 0
run veach --workload veach --res 1280 --spp 256 --cpu-baseline 0
run matball --workload matball --res 1920 --spp 128 --cpu-baseline 0
run grid --workload grid --res 2048 --spp 64 --cpu-baseline 0
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $out/${tag}_bench_reference.json 2> $out/${tag}_bench_reference.err; cut -c1-300 $out/${tag}_bench_reference.json
NCU="ncu --set full --clock-control none --import-source on"
prof() { local name=$1 k=$2 skip=$3; shift 3
  timeout 600 $NCU -k regex:$k -s $skip -c 1 -o /tmp/${name} -f "$@" > $out/${tag}_ncu_${name}.log 2>&1
  ncu -i /tmp/${name}.ncu-rep --page raw --csv > $out/${tag}_ncu_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${name}.ncu-rep --page source --csv > $out/${tag}_ncu_${name}_src.csv 2>/dev/null; rm -f /tmp/${name}.ncu-rep; }
prof cbox_shade k_shade 60 python scripts/render_once.py --workload cbox --spp 16
prof cbox_trace 'k_trace' 60 python scripts/render_once.py --workload cbox --spp 16
prof matball_trace 'k_trace' 40 python scripts/render_once.py --workload matball --res 1920 --spp 16
prof matball_shade 'k_shade' 40 python scripts/render_once.py --workload matball --res 1920 --spp 16
prof grid_trace 'k_trace_inst' 12 python scripts/render_once.py --workload grid --res 1024 --spp 16
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $out/${tag}_launches_cbox.csv python scripts/render_once.py --workload cbox --spp 32 --flags 0 > $out/${tag}_ncu_launches.log 2>&1
rm -f $out/${tag}_ncu_matball_shade_src.csv $out/${tag}_ncu_grid_trace_src.csv; du -sh $out

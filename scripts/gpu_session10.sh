#!/bin/bash
# session 10: k_trace_resume (yield / resume traversal) — parity tests, then A/B against the static kernel on matball / veach
out=gpurun_out; mkdir -p $out; tag=${1:-s10}
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_lbvh_gpu.py -m gpu -q -x -k "yield_resume or bitexact or lbvh or lane_slots or atomics" > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
grep -E "passed|failed|exit|^E  |Error|^FAILED" $out/${tag}_pytest.log | cut -c1-300 | tail -12
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
run() { local name=$1; shift 1; timeout 600 python bench.py --cpu-baseline 0 --steps 3 --warmup 3 "$@" > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err; show $out/${tag}_bench_$name.json $name; }
MB="--workload matball --res 1920 --spp 128"; VE="--workload veach --res 1280 --spp 256"
run matball_static $MB --static-trace
run matball_resume $MB
run matball_thr8 $MB --yield-threshold 8
run matball_thr16 $MB --yield-threshold 16
run matball_thr20 $MB --yield-threshold 20
run veach_static $VE --static-trace
run veach_resume $VE
run veach_thr20 $VE --yield-threshold 20
run matball_static2 $MB --static-trace
NCU="ncu --set full --clock-control none --import-source on"
prof() { local name=$1 k=$2 skip=$3; shift 3
  timeout 600 $NCU -k regex:$k -s $skip -c 1 -o /tmp/${name} -f "$@" > $out/${tag}_ncu_${name}.log 2>&1
  ncu -i /tmp/${name}.ncu-rep --page raw --csv > $out/${tag}_ncu_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${name}.ncu-rep --page source --csv > $out/${tag}_ncu_${name}_src.csv 2>/dev/null; rm -f /tmp/${name}.ncu-rep; }
prof matball_resume 'k_trace' 40 python scripts/render_once.py --workload matball --res 1920 --spp 16
du -sh $out

#!/bin/bash
# session 2: tests (incl. parity at size, LBVH), bench lines, ncu captures, 10 M-triangle LBVH
out=gpurun_out; mkdir -p $out; tag=${1:-s2}
timeout 2400 python -m pytest tests -m gpu -x -q -s > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
grep -E "PARITY|LBVH|passed|failed|exit|Error" $out/${tag}_pytest.log | cut -c1-600 | tail -20
for w in "cbox --bvh-counters" "veach --res 1280 --spp 256" "matball --res 1920 --spp 128" "grid --res 2048 --spp 64"; do
  n=${w%% *}
  timeout 600 python bench.py --cpu-baseline 0 --workload $w > $out/${tag}_bench_$n.json 2> $out/${tag}_bench_$n.err
  python - $out/${tag}_bench_$n.json $n <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = j["roofline"]
    print(f"{sys.argv[2]:8s} {j['value']:8.1f} Msamples/s e2e {j['e2e']['value']:8.1f} step {j['ms_per_step']:7.2f} ms  " +
          "  ".join(f"{k} {v['ms']:.1f} ms {v['frac']:.3f}" for k, v in r["stages"].items()) + f"  step_frac {r['step_frac']:.3f} clk {j['clocks']['sm_mhz']}", j.get("bvh", {}).get("nodes_per_ray"), j.get("bvh", {}).get("tris_per_ray"))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
done
timeout 900 python scripts/lbvh_10m.py > $out/${tag}_lbvh10m.json 2> $out/${tag}_lbvh10m.err; cat $out/${tag}_lbvh10m.json | cut -c1-900
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:k_shade -s 60 -c 1 -o $out/${tag}_ncu_cbox_shade -f python scripts/render_once.py --workload cbox --spp 16 > $out/${tag}_ncu1.log 2>&1
timeout 600 $NCU -k regex:k_trace -s 60 -c 1 -o $out/${tag}_ncu_cbox_trace -f python scripts/render_once.py --workload cbox --spp 16 > $out/${tag}_ncu2.log 2>&1
timeout 600 $NCU -k regex:k_trace -s 40 -c 1 -o $out/${tag}_ncu_matball_trace -f python scripts/render_once.py --workload matball --res 1920 --spp 16 > $out/${tag}_ncu3.log 2>&1
timeout 600 $NCU -k regex:k_trace_inst -s 12 -c 1 -o $out/${tag}_ncu_grid_trace -f python scripts/render_once.py --workload grid --res 1024 --spp 16 > $out/${tag}_ncu4.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file $out/${tag}_launches_cbox.csv python scripts/render_once.py --workload cbox --spp 32 --flags 0 > $out/${tag}_ncu5.log 2>&1
ls -la $out | grep ${tag}_ncu

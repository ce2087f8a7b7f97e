#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=${1:-s5}
timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_parity_at_size.py > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
grep -E "passed|failed|exit|^E  |Error" $out/${tag}_pytest.log | cut -c1-300 | tail -12
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
run() { local name=$1 lib=$2; shift 2; B2PT_LIB=$lib timeout 600 python bench.py --cpu-baseline 0 "$@" > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err; show $out/${tag}_bench_$name.json $name; }
CUR=$PWD/mitsuba2_b200/libb2pt.so; V=$PWD/build_variants
run cbox $CUR --workload cbox --bvh-counters
run cbox_nodefer $V/libb2pt_nodefer.so --workload cbox
run cbox_resgrid $V/libb2pt_resgrid.so --workload cbox
run cbox_again $CUR --workload cbox
run veach $CUR --workload veach --res 1280 --spp 256
run veach_resgrid $V/libb2pt_resgrid.so --workload veach --res 1280 --spp 256
run matball $CUR --workload matball --res 1920 --spp 128
run matball_resgrid $V/libb2pt_resgrid.so --workload matball --res 1920 --spp 128
run grid $CUR --workload grid --res 2048 --spp 64
NCU="ncu --set full --clock-control none --import-source on"
prof() { local name=$1 k=$2 skip=$3; shift 3
  timeout 900 $NCU -k regex:$k -s $skip -c 1 -o /tmp/${name} -f "$@" > $out/${tag}_ncu_${name}.log 2>&1
  ncu -i /tmp/${name}.ncu-rep --page raw --csv > $out/${tag}_ncu_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${name}.ncu-rep --page source --csv > $out/${tag}_ncu_${name}_src.csv 2>/dev/null; rm -f /tmp/${name}.ncu-rep; }
prof cbox_shade k_shade 60 python scripts/render_once.py --workload cbox --spp 16
prof cbox_trace 'k_trace' 60 python scripts/render_once.py --workload cbox --spp 16
du -sh $out

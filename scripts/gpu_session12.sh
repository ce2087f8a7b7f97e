#!/bin/bash
# session 12/13: warp-cooperative tests on the flat leaf list (flat_any_hit_coop, flat_closest_hit_coop) — parity, A/B against the build without the closest-hit part, ncu
out=gpurun_out; mkdir -p $out; tag=${1:-s12}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bitexact or atomics or lane_slots or stage or errors or rank_without" > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
grep -E "passed|failed|exit|^E  |Error|^FAILED" $out/${tag}_pytest.log | cut -c1-300 | tail -12
show() {
  python - $1 $2 <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = j["roofline"]
    print(f"{sys.argv[2]:22s} {j['value']:8.1f} Msamples/s e2e {j['e2e']['value']:8.1f} step {j['ms_per_step']:7.2f} ms  " +
          "  ".join(f"{k} {v['ms']:.1f} ms {v['frac']:.3f} ({v['launches']})" for k, v in r["stages"].items()) + f"  step_frac {r['step_frac']:.3f} clk {j['clocks']['sm_mhz']}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run() { local name=$1 lib=$2; shift 2; B2PT_LIB=$lib timeout 600 python bench.py --cpu-baseline 0 --steps 3 --warmup 3 "$@" > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err; show $out/${tag}_bench_$name.json $name; }
CUR=$PWD/mitsuba2_b200/libb2pt.so; OLD=$PWD/build_variants/libb2pt_noclosest.so
run cbox_coop $CUR
run cbox_nocoop $OLD
run cbox_coop2 $CUR
run cbox_nocoop2 $OLD
run cbox_coop_256 $CUR --trace-block 256 --shade-block 256
run cbox_coop_t64 $CUR --trace-block 64
NCU="ncu --set full --clock-control none --import-source on"
prof() { local name=$1 k=$2 skip=$3; shift 3
  timeout 600 $NCU -k regex:$k -s $skip -c 1 -o /tmp/${name} -f "$@" > $out/${tag}_ncu_${name}.log 2>&1
  ncu -i /tmp/${name}.ncu-rep --page raw --csv > $out/${tag}_ncu_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${name}.ncu-rep --page source --csv > $out/${tag}_ncu_${name}_src.csv 2>/dev/null; rm -f /tmp/${name}.ncu-rep; }
prof cbox_trace 'k_trace' 60 python scripts/render_once.py --workload cbox --spp 16
du -sh $out

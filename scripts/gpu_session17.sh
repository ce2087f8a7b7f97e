#!/bin/bash
# session 17: every hot counter in its own 128-byte line (CNT_ALIVE / CNT_LIST_P / work counter) vs the packed layout
out=gpurun_out; mkdir -p $out; tag=${1:-s15}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bitexact or atomics or lane_slots or stage or yield_resume or c1_reference or progressive" > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
grep -E "passed|failed|exit|^E  |Error|^FAILED" $out/${tag}_pytest.log | cut -c1-300 | tail -8
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
CUR=$PWD/mitsuba2_b200/libb2pt.so; V=$PWD/build_variants/libb2pt_oldcnt.so
MB="--workload matball --res 1920 --spp 128"; VE="--workload veach --res 1280 --spp 256"; GR="--workload grid --res 2048 --spp 64"
run cbox_spread $CUR
run cbox_packed $V
run cbox_spread2 $CUR
run cbox_packed2 $V
run veach_spread $CUR $VE
run veach_packed $V $VE
run matball_spread $CUR $MB
run matball_packed $V $MB
run grid_spread $CUR $GR
run grid_packed $V $GR
NCU="ncu --set full --clock-control none --import-source on"
prof() { local name=$1 k=$2 skip=$3; shift 3
  timeout 600 $NCU -k regex:$k -s $skip -c 1 -o /tmp/${name} -f "$@" > $out/${tag}_ncu_${name}.log 2>&1
  ncu -i /tmp/${name}.ncu-rep --page raw --csv > $out/${tag}_ncu_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${name}.ncu-rep --page source --csv > $out/${tag}_ncu_${name}_src.csv 2>/dev/null; rm -f /tmp/${name}.ncu-rep; }
prof cbox_shade 'k_shade' 60 python scripts/render_once.py --workload cbox --spp 16
du -sh $out

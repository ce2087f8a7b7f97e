#!/bin/bash
# session 18: statistics counters in 8 copies (one 128-byte line each, CTA -> copy blockIdx % 8) vs one line
out=gpurun_out; mkdir -p $out; tag=${1:-s18}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bitexact or atomics or lane_slots or stage or yield_resume or progressive or errors" > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
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
CUR=$PWD/mitsuba2_b200/libb2pt.so; V=$PWD/build_variants/libb2pt_onestat.so
MB="--workload matball --res 1920 --spp 128"; VE="--workload veach --res 1280 --spp 256"; GR="--workload grid --res 2048 --spp 64"
run cbox_stat8 $CUR
run cbox_stat1 $V
run cbox_stat8b $CUR
run cbox_stat1b $V
run veach_stat8 $CUR $VE
run veach_stat1 $V $VE
run grid_stat8 $CUR $GR
run grid_stat1 $V $GR
run matball_stat8 $CUR $MB
du -sh $out

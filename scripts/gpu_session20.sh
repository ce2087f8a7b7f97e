#!/bin/bash
# session 20: flat k_trace loops issue their dependent queue loads up front (A/B against the previous build)
out=gpurun_out; mkdir -p $out; tag=${1:-s20}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bitexact or atomics or lane_slots or stage_shade or errors" > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
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
CUR=$PWD/mitsuba2_b200/libb2pt.so; V=$PWD/build_variants/libb2pt_base.so
run cbox_new $CUR
run cbox_base $V
run cbox_new2 $CUR
run cbox_base2 $V
du -sh $out

#!/bin/bash
# session 11: CTA-size tuning of the two stage kernels (B2_FLAG_TRACE_BLOCK / B2_FLAG_SHADE_BLOCK), k_trace_resume with a wider wavefront
out=gpurun_out; mkdir -p $out; tag=${1:-s11}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "yield_resume or direct_e or veach or matball" > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
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
run() { local name=$1; shift 1; timeout 600 python bench.py --cpu-baseline 0 --steps 3 --warmup 3 "$@" > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err; show $out/${tag}_bench_$name.json $name; }
MB="--workload matball --res 1920 --spp 128"; VE="--workload veach --res 1280 --spp 256"
run cbox_default
run cbox_t128 --trace-block 128
run cbox_s128 --shade-block 128
run cbox_t128_s128 --trace-block 128 --shade-block 128
run cbox_t64_s64 --trace-block 64 --shade-block 64
run cbox_default2
run matball_default $MB
run matball_t128_s128 $MB --trace-block 128 --shade-block 128
run matball_t64_s128 $MB --trace-block 64 --shade-block 128
run matball_resume8_wide $MB --resume-trace --yield-threshold 8 --lane-slots 606208
run matball_resume6 $MB --resume-trace --yield-threshold 6
run veach_t128_s128 $VE --trace-block 128 --shade-block 128
du -sh $out

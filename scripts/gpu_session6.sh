#!/bin/bash
out=gpurun_out; mkdir -p $out; tag=${1:-s6}
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_parity_at_size.py > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
grep -E "passed|failed|exit|^E  |Error|^FAILED" $out/${tag}_pytest.log | cut -c1-300 | tail -12
show() {
  python - $1 $2 <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = j["roofline"]
    print(f"{sys.argv[2]:22s} {j['value']:8.1f} Msamples/s e2e {j['e2e']['value']:8.1f} step {j['ms_per_step']:7.2f} ms  " +
          "  ".join(f"{k} {v['ms']:.1f} ms {v['frac']:.3f}" for k, v in r["stages"].items()) + f"  step_frac {r['step_frac']:.3f} clk {j['clocks']['sm_mhz']}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
run() { local name=$1 lib=$2; shift 2; B2PT_LIB=$lib timeout 600 python bench.py --cpu-baseline 0 "$@" > $out/${tag}_bench_$name.json 2> $out/${tag}_bench_$name.err; show $out/${tag}_bench_$name.json $name; }
CUR=$PWD/mitsuba2_b200/libb2pt.so; V=$PWD/build_variants
run cbox $CUR --workload cbox
for v in stash spec both bothres; do run cbox_$v $V/libb2pt_$v.so --workload cbox; done
run cbox_again $CUR --workload cbox
run veach $CUR --workload veach --res 1280 --spp 256
for v in both bothres; do run veach_$v $V/libb2pt_$v.so --workload veach --res 1280 --spp 256; done
run matball $CUR --workload matball --res 1920 --spp 128
run matball_both $V/libb2pt_both.so --workload matball --res 1920 --spp 128
run grid $CUR --workload grid --res 2048 --spp 64
run grid_ifif $V/libb2pt_ifif.so --workload grid --res 2048 --spp 64
run grid_both $V/libb2pt_both.so --workload grid --res 2048 --spp 64
du -sh $out

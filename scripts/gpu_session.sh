#!/bin/bash
# One gpurun session: GPU tests, then A/B bench lines of every BASELINE workload class (this build vs the
# variants under build_variants/), everything written to gpurun_out/<tag>_*. Usage: scripts/gpu_session.sh <tag> [variants...]
tag=${1:-s}; shift
out=gpurun_out; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $out/${tag}_gpu.txt 2>&1
if [ -z "$SKIP_TESTS" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
  tail -15 $out/${tag}_pytest.log
fi
bench() {   # name lib args...
  local name=$1 lib=$2; shift 2
  B2PT_LIB=$lib timeout 600 python bench.py --cpu-baseline 0 "$@" > $out/${tag}_${name}.json 2> $out/${tag}_${name}.err
  python - "$out/${tag}_${name}.json" "$name" <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    st = j.get("stages") or {}
    print(f"{sys.argv[2]:28s} {j['value']:8.1f} Msamples/s  e2e {j['e2e']['value']:8.1f}  step {j['ms_per_step']:7.2f} ms  " +
          "  ".join(f"{k} {v['ms']:.1f} ms" for k, v in st.items()))
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for v in cur "$@"; do
  lib=$PWD/mitsuba2_b200/libb2pt.so; [ "$v" != cur ] && lib=$PWD/build_variants/libb2pt_$v.so
  [ -f "$lib" ] || { echo "missing $lib"; continue; }
  bench cbox_$v $lib --workload cbox
  bench veach_$v $lib --workload veach --res 1280 --spp 256
  bench matball_$v $lib --workload matball --res 1920 --spp 128
  bench grid_$v $lib --workload grid --res 2048 --spp 64
done

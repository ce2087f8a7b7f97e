#!/bin/bash
# session 14: k_shade at 72 registers / 7 CTAs of 128 threads per SM (variant build) vs 80 / 6; wavefront width with the new traversal kernels
out=gpurun_out; mkdir -p $out; tag=${1:-s14}
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
CUR=$PWD/mitsuba2_b200/libb2pt.so; V=$PWD/build_variants/libb2pt_shade72.so
MB="--workload matball --res 1920 --spp 128"; VE="--workload veach --res 1280 --spp 256"
B2PT_LIB=$V timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "bitexact_single_pass and (cbox_64_spp16-ggx or veach or textures)" 2>&1 | tail -2
run cbox_80 $CUR
run cbox_72 $V
run cbox_80b $CUR
run cbox_72b $V
run cbox_80_slots075 $CUR --lane-slots 340992
run cbox_80_slots125 $CUR --lane-slots 568320
run matball_80_s128 $CUR $MB --shade-block 128
run matball_72_s128 $V $MB --shade-block 128
run veach_80 $CUR $VE
run veach_80_s128 $CUR $VE --shade-block 128
run veach_72_s128 $V $VE --shade-block 128
NCU="ncu --set full --clock-control none --import-source on"
prof() { local name=$1 k=$2 skip=$3; shift 3
  timeout 600 $NCU -k regex:$k -s $skip -c 1 -o /tmp/${name} -f "$@" > $out/${tag}_ncu_${name}.log 2>&1
  ncu -i /tmp/${name}.ncu-rep --page raw --csv > $out/${tag}_ncu_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${name}.ncu-rep --page source --csv > $out/${tag}_ncu_${name}_src.csv 2>/dev/null; rm -f /tmp/${name}.ncu-rep; }
prof cbox_shade 'k_shade' 60 python scripts/render_once.py --workload cbox --spp 16
du -sh $out

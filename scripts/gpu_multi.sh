#!/bin/bash
# gpurun --gpus 8: the north_star configurations at their stated GPU counts (veach-mis class 1280x720 x 1024 spp at 2/4/8,
# instanced lattice 2048^2 x 64 spp at 8), launched the way the driver launches bench.py for N > 1
out=gpurun_out; mkdir -p $out; tag=${1:-m1}
tr() { local n=$1 name=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --cpu-baseline 0 "$@" > $out/${tag}_${name}_n$n.json 2> $out/${tag}_${name}_n$n.err
  python - $out/${tag}_${name}_n$n.json ${name}_n$n <<'PY'
import json, sys
try:
    j = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r = j.get("roofline") or {}
    print(f"{sys.argv[2]:16s} {j['value']:9.1f} Msamples/s e2e {j['e2e']['value']:9.1f} step {j['ms_per_step']:7.2f} ms  " +
          "  ".join(f"{k} {v['ms']:.1f} ms" for k, v in (r.get('stages') or {}).items()) + f"  step_frac {r.get('step_frac')}")
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
for n in 8 2; do timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) scripts/multi_parity.py > $out/${tag}_parity_n$n.json 2> $out/${tag}_parity_n$n.err; tail -c 900 $out/${tag}_parity_n$n.json; tail -3 $out/${tag}_parity_n$n.err | cut -c1-200; done
for n in ${GPUS:-8 4 2}; do tr $n veach --workload veach --res 1280 --spp 1024 --steps 3 --warmup 3; done
tr 8 grid --workload grid --res 2048 --spp 64 --steps 3 --warmup 3
if [ -n "$WITH_CBOX" ]; then for n in 8 4 2; do tr $n cbox --workload cbox --steps 5 --warmup 3; done; fi      # (the driver's own scaling run is exactly this)
# (the 1-GPU line of the same configuration: scripts/gpu_session9.sh)

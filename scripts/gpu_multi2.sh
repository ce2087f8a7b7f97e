#!/bin/bash
# gpurun --gpus 2: sanity of the N-GPU path on the final build — N-GPU film == 1-GPU film through b2_render_reduce + NCCL
# (scripts/multi_parity.py) and the driver's own N = 2 launch of bench.py
out=gpurun_out; mkdir -p $out; tag=${1:-m2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29602 scripts/multi_parity.py > $out/${tag}_parity_n2.json 2> $out/${tag}_parity_n2.err; tail -c 700 $out/${tag}_parity_n2.json; tail -3 $out/${tag}_parity_n2.err | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29502 bench.py --gpus 2 --steps 3 --warmup 3 > $out/${tag}_bench_cbox_n2.json 2> $out/${tag}_bench_cbox_n2.err; cut -c1-400 $out/${tag}_bench_cbox_n2.json; tail -3 $out/${tag}_bench_cbox_n2.err | cut -c1-200

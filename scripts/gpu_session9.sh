#!/bin/bash
# closing session: the whole GPU suite on the final build, configs[2] at stated size through the plugin, veach at 1024 spp on 1 GPU,
# the lattice line + its ncu capture
out=gpurun_out; mkdir -p $out; tag=${1:-s9}
timeout 2400 python -m pytest tests -m gpu -q -s > $out/${tag}_pytest.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest.log
grep -E "passed|failed|exit|^E  |Error|^FAILED" $out/${tag}_pytest.log | cut -c1-300 | tail -12
timeout 600 python scripts/c3_at_size.py 512 > $out/${tag}_c3_at_size.json 2> $out/${tag}_c3_at_size.err; cut -c1-600 $out/${tag}_c3_at_size.json
timeout 600 python bench.py --cpu-baseline 0 --workload veach --res 1280 --spp 1024 > $out/${tag}_bench_veach1024.json 2> $out/${tag}_bench_veach1024.err; cut -c1-200 $out/${tag}_bench_veach1024.json
timeout 600 python bench.py --cpu-baseline 0 --workload grid --res 2048 --spp 64 > $out/${tag}_bench_grid.json 2> $out/${tag}_bench_grid.err; cut -c1-200 $out/${tag}_bench_grid.json
timeout 600 python bench.py --cpu-baseline 0 --workload matball --res 1920 --spp 512 > $out/${tag}_bench_matball512.json 2> $out/${tag}_bench_matball512.err; cut -c1-200 $out/${tag}_bench_matball512.json
NCU="ncu --set full --clock-control none --import-source on"
timeout 900 $NCU -k regex:k_trace_inst -s 12 -c 1 -o /tmp/grid_trace -f python scripts/render_once.py --workload grid --res 1024 --spp 16 > $out/${tag}_ncu_grid_trace.log 2>&1
ncu -i /tmp/grid_trace.ncu-rep --page raw --csv > $out/${tag}_ncu_grid_trace_raw.csv 2>/dev/null; rm -f /tmp/grid_trace.ncu-rep
du -sh $out

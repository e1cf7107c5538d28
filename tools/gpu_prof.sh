#!/bin/bash
# ncu evidence for profiles/: depth-wise attention (full set), launch list of one bench step
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
python tools/dw_one.py 40 15 16 20 276 && timeout 600 ncu --set full --import-source on --clock-control none -k regex:dw_attn -c 2 -s 4 -o gpurun_out/dw_attn_n15 python tools/dw_one.py 40 15 16 20 276 > gpurun_out/ncu_dw.log 2>&1; tail -2 gpurun_out/ncu_dw.log
timeout 600 ncu --set full --clock-control none -k regex:dw_attn -c 2 -s 4 -o gpurun_out/dw_attn_n4 python tools/dw_one.py 40 4 16 20 276 > gpurun_out/ncu_dw4.log 2>&1; tail -1 gpurun_out/ncu_dw4.log
timeout 600 python bench.py --steps 2 --warmup 3 --eager --no-roofline --no-cpu-baseline > gpurun_out/bench_eager.json 2>/dev/null && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 1 --warmup 3 --eager --no-roofline --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?"; wc -l gpurun_out/launches_bench.csv

#!/bin/bash
# ncu evidence for profiles/: nonlocal kernels (full set) and the launch list of ONE eager training step
mkdir -p gpurun_out
python tools/nl_one.py 40 16 32 276 --bwd > /dev/null && timeout 600 ncu --set full --import-source on --clock-control none -k regex:nl_ -c 3 -s 6 -o gpurun_out/nl_tc_v3 python tools/nl_one.py 40 16 32 276 --bwd > gpurun_out/ncu_nl3.log 2>&1; tail -1 gpurun_out/ncu_nl3.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv -s 22000 -c 11000 --log-file gpurun_out/launches_step.csv python bench.py --steps 1 --warmup 3 --eager --no-roofline --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench rc=$?"; wc -l gpurun_out/launches_step.csv

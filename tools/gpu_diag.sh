#!/bin/bash
# Round diagnostics on the GPU box: parity tests, bench (graph), step profile.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_graph.json 2> gpurun_out/bench_graph.err; echo "bench graph rc=$?"
timeout 600 python tools/step_probe.py > gpurun_out/step_probe.txt 2>&1; echo "probe rc=$?"
cut -c1-700 gpurun_out/bench_graph.json

#!/bin/bash
# One gpurun call for a nonlocal-kernel iteration (~1.3 GPU-min): parity + timings of every shape, the nonlocal
# GPU tests, and the phase timeline of one CTA per role.  Build both libraries HERE first:
#   python -m deep_attentive_vi_b200.build && python tools/nl_trace.py --build-only
mkdir -p gpurun_out
timeout 60 python tools/nl_check.py --bwd > gpurun_out/nl_check.jsonl 2>&1; cut -c1-420 gpurun_out/nl_check.jsonl | tail -9
timeout 90 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k nonlocal 2>&1 | tail -2
timeout 40 python tools/nl_trace.py 40 16 32 276 > gpurun_out/nl_trace.txt 2>&1; tail -5 gpurun_out/nl_trace.txt

#!/bin/bash
# one-call depth-wise iteration: parity tests of the dw kernels + the step's launches replayed
tag=${1:-x}
python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "dw_attn" > gpurun_out/dw_${tag}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/dw_${tag}_pytest.log
python tools/dw_replay.py --trace profiles/r02_dw_trace_cifar16_b40.json > gpurun_out/dw_${tag}_replay.json 2> gpurun_out/dw_${tag}_replay.err
tail -2 gpurun_out/dw_${tag}_pytest.log; tail -c 300 gpurun_out/dw_${tag}_replay.err

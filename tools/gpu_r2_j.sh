#!/bin/bash
# Round 2, GPU call J: the general-shape chain (new tests first), then the whole GPU suite and the smoke entry
mkdir -p gpurun_out
timeout -k 5 300 python -m pytest tests/test_gpu_parity.py -x -q -s -k "free_shape or ref_shapes" > gpurun_out/j_t1.log 2>&1; echo "t1 rc=$?"
grep -E "max\|delta\||passed|failed|Error|error" gpurun_out/j_t1.log | tail -25
timeout -k 5 1200 python -m pytest tests -q -m gpu > gpurun_out/j_t2.log 2>&1; echo "t2 rc=$?"
tail -8 gpurun_out/j_t2.log
timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/j_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/j_smoke.log

#!/bin/bash
# 512-thread folded form of k_flow_tc: bit-identity for every segmentation, then the launch-form sweep over the job size.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout -k 5 240 python -m pytest tests/test_gpu_parity.py -q -x -k "flow_kernel" > gpurun_out/t_flow7.log 2>&1; RC=$?; echo "flow tests rc=$RC"; tail -12 gpurun_out/t_flow7.log
if [ $RC -ne 0 ]; then exit 1; fi
timeout -k 5 300 python tools/sweep_modes.py 1,2,4,8,12,16,32,64 16000 2>gpurun_out/sweep_modes2.err | tee gpurun_out/sweep_modes2.jsonl
PWV_TC_SEG=100 PWV_TRACE_FLOW=1 timeout -k 5 100 python tools/tc_trace.py f16x3 2 > gpurun_out/trace_flow_folded_f16x3_l2.txt 2>&1
head -10 gpurun_out/trace_flow_folded_f16x3_l2.txt | cut -c1-330

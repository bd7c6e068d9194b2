timeout 120 python tools/tc_trace.py f16x3 2>&1 | tee gpurun_out/tc_trace_v3.txt

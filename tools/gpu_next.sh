#!/bin/bash
# First GPU call of the next round (prepared, not yet run): the experimental 512-thread form of k_flow_tc
# (PWV_TC_QUIET=1: helper work folded into the slots' head warps, 128 registers per thread).
#  1. where does it stop? one utterance (tiles per CTA in {1, 2}), one launch per layer, under a short timeout and under
#     compute-sanitizer synccheck (the sweep of round 1 did not get past this size);
#  2. if it runs: its one-launch-per-layer mode (no flags, no fences) against k_layer_tc over the job size.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PWV_TC_QUIET=1
for seg in 1 100; do
  PWV_TC_SEG=$seg timeout -k 5 40 python tools/sweep_modes.py 1 16000 > gpurun_out/next_folded_n1_seg$seg.jsonl 2> gpurun_out/next_folded_n1_seg$seg.err
  echo "folded form, N=1, seg=$seg: rc=$? (124 = hang)"; cat gpurun_out/next_folded_n1_seg$seg.jsonl
done
PWV_TC_SEG=1 timeout -k 5 200 compute-sanitizer --tool synccheck python tools/sanitize.py > gpurun_out/next_folded_synccheck.txt 2>&1; echo "synccheck rc=$?"; tail -5 gpurun_out/next_folded_synccheck.txt
timeout -k 5 200 python tools/sweep_modes.py 2,8,16,64 16000 2> gpurun_out/next_sweep.err | tee gpurun_out/next_sweep.jsonl

#!/bin/bash
# compute-sanitizer memcheck / racecheck / synccheck over the stress script (every named shape at
# three block sizes through huf_encode / huf_decode, i.e. the host lanes, all kernels, both decode
# lanes).  Usage: gpurun --timeout 1800 -- 'bash scripts/gpu_sanitize.sh [mib]'
mkdir -p gpurun_out
MIB=${1:-2}
for tool in memcheck racecheck synccheck; do
    echo "== $tool"
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/stress_gpu.py $MIB > gpurun_out/r2_san_$tool.log 2>&1
    echo "$tool rc=$?"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|failures:" gpurun_out/r2_san_$tool.log | tail -3
done

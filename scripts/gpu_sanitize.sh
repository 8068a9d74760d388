#!/bin/bash
# compute-sanitizer memcheck / racecheck / synccheck over the stress script (every named shape at
# three block sizes through huf_encode / huf_decode, i.e. the host lanes, all kernels, both decode
# lanes), and memcheck once more with the encoder's pass pipeline forced on for small inputs
# (passes over workspace slots and side streams).  Usage: gpurun --timeout 1800 -- 'bash scripts/gpu_sanitize.sh [mib]'
mkdir -p gpurun_out
MIB=${1:-2}
for tool in memcheck racecheck synccheck; do
    echo "== $tool"
    timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python scripts/stress_gpu.py $MIB > gpurun_out/r2_san_$tool.log 2>&1
    echo "$tool rc=$?"
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|failures:" gpurun_out/r2_san_$tool.log | tail -3
done
echo "== memcheck, pass pipeline forced"
HUF_B200_ENC_PIPE_MIN=1 HUF_B200_ENC_PIPE_PASS=1 HUF_B200_ENC_SLOTS=3 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 \
    python scripts/stress_gpu.py $MIB > gpurun_out/r2_san_memcheck_pipeline.log 2>&1
echo "memcheck(pipeline) rc=$?"
grep -E "ERROR SUMMARY|failures:" gpurun_out/r2_san_memcheck_pipeline.log | tail -3

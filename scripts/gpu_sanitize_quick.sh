#!/bin/bash
# memcheck over every shape of scripts/stress_gpu.py (1 MiB each) and racecheck over the three that reach every packing lane, sized for ~3.5 minutes of box time.
# Usage: gpurun --timeout 440 -- "bash scripts/gpu_sanitize_quick.sh"
mkdir -p gpurun_out
echo "== memcheck"
timeout 200 compute-sanitizer --tool memcheck --error-exitcode 7 python scripts/stress_gpu.py 1 > gpurun_out/r2d_san_memcheck.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|failures:" gpurun_out/r2d_san_memcheck.log | tail -3
echo "== racecheck"
HUF_STRESS_SHAPES=zipf255,fibonacci,fib1m timeout 190 compute-sanitizer --tool racecheck --error-exitcode 7 python scripts/stress_gpu.py 1 > gpurun_out/r2d_san_racecheck.log 2>&1; echo "racecheck rc=$?"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|failures:" gpurun_out/r2d_san_racecheck.log | tail -3

#!/bin/bash
# ncu --set full capture of the hot kernels (one launch each) + launch list. One GPU only.
mkdir -p gpurun_out
KREGEX=${1:-"k_pack|k_build|k_seg_hist|k_decode|k_find"}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$KREGEX" -c 6 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline > gpurun_out/prof.log 2>&1
echo "ncu full rc=$?"; tail -3 gpurun_out/prof.log; ls -la gpurun_out/*.ncu-rep

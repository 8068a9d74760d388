#!/bin/bash
# ncu --set full capture of the decode kernels (skips the encode launches of the warm-up). One GPU only.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_find|k_tree|k_decode" -c 4 -f -o gpurun_out/r2_prof_dec \
    python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-extra --no-strong > gpurun_out/r2_prof_dec.log 2>&1
echo "ncu full rc=$?"; tail -2 gpurun_out/r2_prof_dec.log | cut -c1-300; ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# One gpurun call that produces everything scripts/make_profiles.py turns into profiles/<tag>_*:
# smoke, the gpu test lane, the default bench line, the ncu launch list of the same command, and
# ncu --set full captures of the encode and decode kernels.  One GPU only.
# Usage: gpurun --timeout 2400 -- 'bash scripts/gpu_capture.sh'
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1200 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-extra --no-strong > gpurun_out/ncu_bench.log 2>&1; echo "ncu rc=$?"
echo "== ncu full encode"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_pack|k_build|k_seg_hist" -c 6 -f -o gpurun_out/prof \
    python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-extra --no-strong > gpurun_out/prof.log 2>&1; echo "rc=$?"
echo "== ncu full decode"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_find|k_tree|k_decode" -c 4 -f -o gpurun_out/prof_dec \
    python bench.py --steps 1 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-extra --no-strong > gpurun_out/prof_dec.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep

#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): regenerates the evidence under profiles/ into gpurun_out/r01/.
# usage: gpurun --timeout 1500 -- 'bash tools/refresh_profiles.sh'
set -u
O=gpurun_out/r01; mkdir -p $O
PY=python
# 1. bench line (no profiler)
timeout 400 $PY bench.py --steps 50 --warmup 5 > $O/r01_bench_n1.json 2> $O/bench_n1.err
# 2. launch list of the bench command
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 20 -c 160 --csv --log-file $O/r01_launches_bench.csv \
    $PY bench.py --steps 10 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
# 3. the dominant kernel, full set (config 2 size and the 32768-chunk steady state)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_build_fused -s 3 -c 1 -f -o $O/fused_2048 \
    $PY tools/scratch/fused_once.py 8 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_build_fused -s 3 -c 1 -f -o $O/fused_32768 \
    $PY tools/scratch/fused_once.py 32 > /dev/null 2>&1
# 4. staged extraction kernels at 32768 chunks, large-chunk kernels at config 4
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_classify_spec|k_scan_chunks|k_emit_small" -s 6 -c 3 -f -o $O/staged_32768 \
    $PY tools/scratch/staged_once.py 32 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_noise_big|k_count_big|k_emit_big" -s 4 -c 4 -f -o $O/config4 \
    $PY tools/scratch/config4_once.py > /dev/null 2>&1
# 5. other configs
timeout 300 $PY tools/scratch/large.py > $O/config3.txt 2>&1
timeout 300 $PY tools/scratch/config4.py > $O/config4.txt 2>&1
timeout 300 $PY tools/scratch/staged.py > $O/staged_stages.txt 2>&1
timeout 300 $PY tools/scratch/e2e_pipe.py > $O/e2e_pipe.txt 2>&1
timeout 600 $PY tools/bench_flythrough.py --cpu > $O/r01_flythrough_config5.json 2> $O/fly.err
ls -la $O

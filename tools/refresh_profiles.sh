#!/bin/bash
# Runs ON THE GPU BOX (via gpurun): regenerates the round's evidence into gpurun_out/r02/; copy what should be judged
# into profiles/ (see profiles/README.md).   usage: gpurun --timeout 1500 -- 'bash tools/refresh_profiles.sh'
set -u
O=gpurun_out/r02; mkdir -p $O
PY=python
NCU="ncu --set full --clock-control none --import-source on -f"
# 1. bench line (no profiler)
timeout 600 $PY bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/bench_n1.err
# 2. launch list of the bench command (per-launch device time of every kernel of a short run)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv \
    $PY bench.py --steps 3 --warmup 3 --quick --no-cpu-baseline > /dev/null 2>&1
# 3. the dominant kernel, full set: the bench workload (config 3), config 2 and the 32768-chunk steady state
timeout 600 $NCU -k regex:k_build_fused -s 2 -c 1 -o $O/fused_config3 $PY tools/once.py config3 > /dev/null 2>&1
timeout 400 $NCU -k regex:k_build_fused -s 3 -c 1 -o $O/fused_2048 $PY tools/once.py fused 8 > /dev/null 2>&1
timeout 400 $NCU -k regex:k_build_fused -s 3 -c 1 -o $O/fused_32768 $PY tools/once.py fused 32 > /dev/null 2>&1
# 4. staged pipeline at 32768 chunks (noise stage included), large-chunk kernels at config 4
timeout 400 $NCU -k regex:"k_noise_spec|k_classify_spec|k_scan_chunks|k_emit_small" -s 8 -c 4 -o $O/staged_32768 $PY tools/once.py staged 32 > /dev/null 2>&1
timeout 600 $NCU -k regex:"k_noise_big|k_count_big|k_emit_big" -s 4 -c 4 -o $O/config4 $PY tools/once.py config4 > /dev/null 2>&1
ls -la $O

set -u
O=gpurun_out/r02; mkdir -p $O
PY=python
NCU="ncu --set full --clock-control none --import-source on -f"
$PY -m pytest tests -m gpu -q -s > $O/r02_pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/r02_pytest_gpu.txt
timeout 900 $PY bench.py --steps 20 --warmup 5 > $O/r02_bench_full_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"
timeout 600 $PY bench.py --impl reference --steps 20 --warmup 5 > $O/r02_bench_reference_n1.json 2> $O/bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches_bench.csv \
    $PY bench.py --steps 3 --warmup 3 --quick --no-cpu-baseline > /dev/null 2>&1; echo "launch list rc=$?"
timeout 600 $NCU -k regex:k_build_fused -s 2 -c 1 -o $O/fused_config3 $PY tools/once.py config3 > /dev/null 2>&1; echo "ncu c3 rc=$?"
timeout 400 $NCU -k regex:k_build_fused -s 3 -c 1 -o $O/fused_2048 $PY tools/once.py fused 8 > /dev/null 2>&1; echo "ncu c2 rc=$?"
timeout 400 $NCU -k regex:"k_noise_spec|k_classify_spec|k_scan_chunks|k_emit_small" -s 8 -c 4 -o $O/staged_32768 $PY tools/once.py staged 32 > /dev/null 2>&1; echo "ncu staged rc=$?"
(timeout 900 compute-sanitizer --tool memcheck $PY tools/sanity_paths.py 2>&1 | tail -25; timeout 900 compute-sanitizer --tool racecheck $PY tools/sanity_paths.py 2>&1 | tail -12) > $O/r02_sanitizer.txt; tail -3 $O/r02_sanitizer.txt
ls -la $O

#!/bin/bash
# usage (on the GPU box): bash tools/scratch/abrun.sh <variantA> <variantB> ...   (names under lib/variants, or "main")
# runs the GPU tests on main, then A/B timings: config 2 (flushed, per launch), config 3 (one launch), staged stages
O=gpurun_out/ab; mkdir -p $O
V=underwaterworld_b200/lib/variants
libs=""
for n in "$@"; do if [ "$n" = main ]; then libs="$libs underwaterworld_b200/lib/libuwcuda.so"; else libs="$libs $V/$n.so"; fi; done
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/pytest.log
tail -3 $O/pytest.log
timeout 300 python tools/scratch/ab.py $libs 2>&1 | tee $O/ab.txt
timeout 300 python tools/scratch/large_ab.py $libs 2>&1 | tee $O/large_ab.txt
timeout 200 python tools/scratch/staged.py 2>&1 | tee $O/staged.txt
timeout 300 python tools/scratch/config4.py 2>&1 | tail -5 | tee $O/config4.txt

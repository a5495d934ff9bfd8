#!/bin/bash
# usage (GPU box): bash tools/scratch/abfine.sh <variant> ...   GPU tests on main, then fine-grained config-2 timing + config 3 + staged emit per variant
O=gpurun_out/ab; mkdir -p $O
V=underwaterworld_b200/lib/variants
libs=""
for n in "$@"; do libs="$libs $V/$n.so"; done
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > $O/pytest.log; tail -2 $O/pytest.log
timeout 600 python tools/scratch/ab_fine.py $libs 2>&1 | tee $O/ab_fine.txt
timeout 300 python tools/scratch/large_ab.py $libs 2>&1 | tee $O/large_ab.txt
for n in "$@"; do echo "staged $n"; UWCUDA_LIB=$PWD/$V/$n.so timeout 200 python tools/scratch/staged.py 2>&1 | tail -1 | cut -c1-120; done | tee $O/staged_v.txt

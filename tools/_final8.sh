set -u
O=gpurun_out/r02s; mkdir -p $O
nvidia-smi topo -m > $O/topo.txt 2>&1
run() { N=$1; TAG=$2
  timeout 600 env $3 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 20 --warmup 5 > $O/bench_n${N}_$TAG.json 2> $O/bench_n${N}_$TAG.err; echo "bench N=$N $TAG rc=$?"
}
run 8 staged UW_X=1
run 8 direct UW_STAGED_STORES=0
run 4 staged UW_X=1
run 4 direct UW_STAGED_STORES=0
run 2 staged UW_X=1
timeout 600 python bench.py --steps 20 --warmup 5 --quick --no-cpu-baseline > $O/bench_n1_staged.json 2> $O/bench_n1.err; echo "bench1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29540 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 > $O/bench_ref_n8.json 2> $O/bench_ref_n8.err; echo "ref8 rc=$?"
python -m pytest tests/test_gpu_gather.py -m gpu -x -q > $O/pytest_gather.log 2>&1; echo "pytest gather rc=$?"; tail -3 $O/pytest_gather.log

#!/bin/bash
# Multi-GPU validation on ONE box (gpurun --gpus N): 2-rank parity test, then bench lines at N = 2 ... $1 for C2 (and C4 at the top N).
N=${1:-2}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ddp.py -m gpu -q 2>&1 | tail -3 | tee $O/r02_pytest_gpu_2gpu.log
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 50 --warmup 5 \
      > $O/r02_bench_n$n.json 2> $O/r02_bench_n$n.err
  python tools/bench_line.py c2.n$n $O/r02_bench_n$n.json
  grep -i -m3 "NVLS\|nranks\|Connected all" $O/r02_bench_n$n.err | cut -c1-200
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config c4 --steps 20 --warmup 3 \
    > $O/r02_bench_c4_n$N.json 2> $O/r02_bench_c4_n$N.err
python tools/bench_line.py c4.n$N $O/r02_bench_c4_n$N.json

#!/bin/bash
# quick GPU regression: conv / wgrad tests, then C4 + C2 bench lines (+ optional A/B env in $1, e.g. SHOTVAE_HALO=0)
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "wgrad or conv_fprop or convT or bn_backward" 2>&1 | tail -3
for c in c4 c2; do
  timeout 250 python bench.py --config $c --steps 10 --warmup 3 --cpu-steps 0 --gpu-reference-steps 0 --dump-kernels gpurun_out/r02_kernels_${c}_x.json > gpurun_out/r02_bench_${c}_x.json 2> gpurun_out/r02_bench_${c}_x.err
  python tools/bench_line.py $c gpurun_out/r02_bench_${c}_x.json
done
if [ -n "$1" ]; then
  env $1 timeout 250 python bench.py --steps 20 --warmup 3 --cpu-steps 0 --gpu-reference-steps 0 --dump-kernels gpurun_out/r02_kernels_c2_ab.json > gpurun_out/r02_bench_c2_ab.json 2> gpurun_out/r02_bench_c2_ab.err
  python tools/bench_line.py "$1" gpurun_out/r02_bench_c2_ab.json
fi

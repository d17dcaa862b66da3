#!/bin/bash
# Round-2 full GPU validation on ONE B200: GPU tests, smoke, bench lines (C2 default, C3, C4, C5, --kernels), the ncu launch
# list of three steady-state steps and `ncu --set full` captures of the dominant kernels.  Everything lands in gpurun_out/.
# usage: tools/r02_full_check.sh [tag]      (tag defaults to r02)
T=${1:-r02}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/${T}_clocks.csv &
SMI=$!
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > $O/${T}_pytest_gpu.log
tail -3 $O/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${T}_smoke.log 2>&1; tail -2 $O/${T}_smoke.log
timeout 600 python bench.py --dump-kernels $O/${T}_kernels_c2.json > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
python tools/bench_line.py c2 $O/${T}_bench_n1.json
for c in c3 c4 c5; do
  timeout 600 python bench.py --config $c --steps 20 --warmup 3 --cpu-steps 2 --gpu-reference-steps 10 --dump-kernels $O/${T}_kernels_$c.json > $O/${T}_bench_$c.json 2> $O/${T}_bench_$c.err
  python tools/bench_line.py $c $O/${T}_bench_$c.json
done
timeout 300 python bench.py --kernels > $O/${T}_bandwidth_kernels.json 2> $O/${T}_bandwidth_kernels.err
# kernel timeline of one CUDA-graph replay (CUPTI through torch.profiler): what runs beside what, gaps on the critical chain
for c in c2 c4; do timeout 300 python tools/step_timeline.py $c ${T}_$c > $O/${T}_timeline_$c.txt 2>&1; done
kill $SMI
if [ -z "$NO_NCU" ]; then
  # launch list: eager (--no-graph) so that launch order = program order; skip the build-up steps
  L=$(python -c "import json;print(json.load(open('$O/${T}_bench_n1.json'))['details']['launches_per_step'])")
  timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --launch-skip $((L*3+60)) -c $((L*3)) \
      --csv --log-file $O/${T}_launches_n1.csv python bench.py --steps 3 --warmup 3 --cpu-steps 0 --gpu-reference-steps 0 --no-graph > $O/${T}_ncu_list.log 2>&1
  python tools/launch_summary.py $O/${T}_launches_n1.csv $L > $O/${T}_launches_n1_summary.md 2>&1
  for k in igemm_halo_kernel wgrad_halo_kernel igemm_fprop_tc; do
    timeout 600 ncu --set full --import-source on --clock-control none -k regex:$k --launch-skip 40 -c 3 -f -o $O/${T}_${k}_full \
        python bench.py --steps 1 --warmup 1 --cpu-steps 0 --gpu-reference-steps 0 --no-graph > $O/${T}_ncu_$k.log 2>&1
  done
fi
ls -la $O | tail -40

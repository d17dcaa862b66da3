#!/bin/bash
# `ncu --set full` captures of the kernels that dominate the C2 step (two launches each, eager replay of bench.py): the block-1
# forward conv, the block-1 input-gradient conv with the fused BatchNorm-backward statistics, the block-1 weight gradient and
# bn_bwd_apply.  Reports land in gpurun_out/<tag>_<name>.ncu-rep; summarise with tools/ncu_pipe_summary.py.
T=${1:-r02t}
O=gpurun_out
mkdir -p $O
i=0
for k in 'igemm_halo_kernel<\(int\)9, \(int\)2, \(bool\)0>' 'igemm_halo_kernel<\(int\)9, \(int\)2, \(bool\)1>' 'wgrad_halo_kernel<\(int\)9>' 'bn_bwd_apply_kernel<\(int\)1'; do
  n=$(echo halo920 halo921 wgrad9 bnapply1 | cut -d' ' -f$((i+1))); i=$((i+1))
  timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"$k" --launch-skip 4 -c 2 -f -o $O/${T}_$n \
      python bench.py --steps 1 --warmup 1 --cpu-steps 0 --gpu-reference-steps 0 --no-graph > $O/${T}_ncu_$n.log 2>&1
  grep -c "igemm_halo\|wgrad_halo\|bn_bwd_apply" $O/${T}_ncu_$n.log; grep "No kernels" $O/${T}_ncu_$n.log
done
ls -la $O/${T}_*.ncu-rep

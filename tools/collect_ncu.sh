#!/bin/bash
# Round profile collection on the GPU box (run through gpurun): ncu --set full on one mid-solve Krylov
# iteration of each workload plus the constraint-stage kernels; the reports stay in /tmp on the box,
# only their raw-metric CSV pages come back (the .ncu-rep files exceed gpurun's 64 MiB return limit).
set -u
OUT=gpurun_out
mkdir -p $OUT
cap() {  # name, extra env, ncu args...
  local name=$1; shift
  local envs=$1; shift
  env $envs ncu --set full --clock-control none "$@" -o /tmp/$name -f python tools/ncu_target.py 10000000 1 > $OUT/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $OUT/ncu_full_$name.csv 2>> $OUT/ncu_$name.log
}
cap r1_lkdv_iter "SPIS_WORKLOAD=lkdv" -k regex:"spmv_sell|mdot_kernel|lincomb|orth_mid|scale_kernel" --launch-skip 118 --launch-count 8
cap r1_lkdv_mdotm "SPIS_WORKLOAD=lkdv" -k regex:"mdotm" --launch-count 2
cap r1_swe_spmv "SPIS_WORKLOAD=swe" -k regex:"spmv_sell" --launch-skip 20 --launch-count 3
ls -la /tmp/*.ncu-rep >> $OUT/ncu_sizes.log

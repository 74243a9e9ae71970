#!/bin/bash
# Round-2 profile collection on the GPU box (run through gpurun, ONE GPU): launch lists (time + DRAM bytes of every
# kernel of one solve) and ncu --set full on one mid-solve Krylov iteration of each workload, the constraint-stage
# kernels and the preconditioner kernels; the reports stay in /tmp on the box, only CSV pages come back (the .ncu-rep
# files exceed gpurun's 64 MiB return limit).
set -u
OUT=gpurun_out
R=${ROUND:-r2}
mkdir -p $OUT
for WL in lkdv swe; do
  SPIS_WORKLOAD=$WL python tools/ncu_target.py 10000000 2 > $OUT/plain_$WL.log 2>&1 || exit 1
  SPIS_WORKLOAD=$WL ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
      --log-file $OUT/launches_${R}_$WL.csv python tools/ncu_target.py 10000000 1 > $OUT/ncu_list_$WL.log 2>&1
done
cap() {  # name, env, size, ncu args...
  local name=$1; shift
  local envs=$1; shift
  local size=$1; shift
  env $envs ncu --set full --import-source on --clock-control none "$@" -o /tmp/$name -f python tools/ncu_target.py $size 1 > $OUT/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $OUT/ncu_full_$name.csv 2>> $OUT/ncu_$name.log
}
# one mid-solve iteration of the pipelined loop: dual SpMV (field windows), mdot with the riding residual, orth_mid with
# the norm, hess_kernel, lincomb2n (5 kernels per step; step 15 starts near launch 3 + 15 * 5)
cap ${R}_lkdv_iter "SPIS_WORKLOAD=lkdv" 10000000 -k regex:"spmv_|mdot_|lincomb|orth_mid|hess_kernel" --launch-skip 78 --launch-count 10
cap ${R}_lkdv_gram "SPIS_WORKLOAD=lkdv" 10000000 -k regex:"gram_kernel|mdotm|spmv_pattern_multi" --launch-count 7
cap ${R}_swe_iter "SPIS_WORKLOAD=swe" 10000000 -k regex:"spmv_|mdot_|lincomb|orth_mid|hess_kernel" --launch-skip 53 --launch-count 6
cap ${R}_pre_jacobi "SPIS_WORKLOAD=jacobi" 10000000 -k regex:"lincomb2n|lincomb_kernel|jacobi" --launch-skip 4 --launch-count 4
cap ${R}_pre_block "SPIS_WORKLOAD=lkdvRK" 6000000 -k regex:"blockdiag|lincomb_kernel" --launch-skip 2 --launch-count 3
cap ${R}_pre_csr "SPIS_WORKLOAD=csrpre" 10000000 -k regex:"spmv_" --launch-skip 4 --launch-count 3
ls -la /tmp/*.ncu-rep >> $OUT/ncu_sizes.log

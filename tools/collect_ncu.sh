#!/bin/bash
# Round profile collection on the GPU box (run through gpurun): launch lists (time + DRAM bytes of every
# kernel of one solve) and ncu --set full on one mid-solve Krylov iteration of each workload plus the
# constraint-stage kernels; the reports stay in /tmp on the box, only CSV pages come back (the .ncu-rep
# files exceed gpurun's 64 MiB return limit).
set -u
OUT=gpurun_out
mkdir -p $OUT
for WL in lkdv swe; do
  SPIS_WORKLOAD=$WL python tools/ncu_target.py 10000000 2 > $OUT/plain_$WL.log 2>&1 || exit 1
  SPIS_WORKLOAD=$WL ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv \
      --log-file $OUT/launches_r1_$WL.csv python tools/ncu_target.py 10000000 1 > $OUT/ncu_list_$WL.log 2>&1
done
cap() {  # name, env, ncu args...
  local name=$1; shift
  local envs=$1; shift
  env $envs ncu --set full --clock-control none "$@" -o /tmp/$name -f python tools/ncu_target.py 10000000 1 > $OUT/ncu_$name.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $OUT/ncu_full_$name.csv 2>> $OUT/ncu_$name.log
}
cap r1_lkdv_iter "SPIS_WORKLOAD=lkdv" -k regex:"spmv_|mdot_|lincomb|orth_mid|scale_kernel|reduce_partials" --launch-skip 96 --launch-count 8
cap r1_lkdv_mdotm "SPIS_WORKLOAD=lkdv" -k regex:"mdotm" --launch-count 2
cap r1_swe_spmv "SPIS_WORKLOAD=swe" -k regex:"spmv_sell_dual|spmv_selld" --launch-skip 8 --launch-count 3
ls -la /tmp/*.ncu-rep >> $OUT/ncu_sizes.log

#!/bin/bash
# SASS of the hot kernels out of the built library -> profiles/sass/<kernel>.sass (cuobjdump, no GPU needed)
set -eu
LIB=structurepreservingiterativesolvers_b200/lib/libspis_b200.so
OUT=profiles/sass
mkdir -p $OUT
cuobjdump -sass $LIB > /tmp/spis_all.sass
for pat in 'spmv_pattern_kernelILi0ELi2ELb1' 'spmv_pattern_dual_kernelILi2' 'spmv_sell_dual_kernelILb1' 'spmv_sell_dual_kernelILb0' \
           'spmv_selld_kernelILi0' 'spmv_selld_kernelILi2' 'spmv_sellp_kernelILi0ELb0' 'spmv_sell_kernelILi0' \
           'mdot_kernelILi2' 'mdot_kernelILi4' 'mdot_reg_kernelILi24' 'mdotm_kernelILi4' 'spmv_pattern_multi_kernelILi2ELi4' 'spmv_sell_multi_kernelILb1ELi4' 'lincomb_kernelILi4' 'lincomb2_kernelILi4' \
           'orth_mid_kernelILi24ELi2' 'scale_kernel' 'jacobi_kernel' 'blockdiag_kernelILi3' \
           'lincomb2n_kernelILi4' 'hess_kernel' 'spmv_fw_kernelILi2ELi3ELi2ELi8' 'spmv_fw_kernelILi1ELi0ELi2ELi8' 'spmv_sellw_kernelILi1ELi0ELb1' \
           'gram_kernelILi3' 'gram_kernelILi7' 'halo_xchg_kernel' 'xreduce_kernel' 'spmv_sell2_kernelILi0' 'spmv_csr_kernelILi8ELi0' 'publish_res_kernel' 'pipe_init_kernel' \
           'mdot_reg_kernelILi8' 'blockdiag_kernelILi6'; do
  awk -v pat="$pat" '/Function : /{f = index($0, pat) > 0} f' /tmp/spis_all.sass > $OUT/$pat.sass
  echo "$pat: $(grep -c ';' $OUT/$pat.sass) instructions"
done

#!/bin/bash
# Developer A/B under ncu: instruction and L1 wavefront counts of k_sim for the in-tree build and every _ab/*.so
for lib in bourse_b200/libbourse_b200.so _ab/*.so; do
  [ -f "$lib" ] || continue
  BOURSE_B200_LIB=$PWD/$lib PYTHONPATH=. ncu --metrics smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_local_op_st.sum --clock-control none -k regex:k_sim -c 1 --csv python scripts/dbg_dense.py 4096 100 dense 2>/dev/null | grep k_sim | awk -F'","' -v l=$lib '{print l, $(NF-2), $NF}'
done

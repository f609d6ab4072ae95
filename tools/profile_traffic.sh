#!/bin/bash
# DRAM / L2 / local-memory traffic of the batch filter kernel under different output sets and P
M=dram__bytes_write.sum,dram__bytes_read.sum,lts__t_bytes.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_st.sum,l1tex__t_bytes_pipe_lsu_mem_local_op_ld.sum,l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum,l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,gpu__time_duration.sum
run() { # name flags extra-env...
  name=$1; shift
  env "$@" ncu --metrics $M --clock-control none -k regex:filter_kernel -c 1 --csv --log-file gpurun_out/traffic_$name.csv python tools/profile_run.py 18 1 1000 > /dev/null 2>&1
}
run all PROF_FLAGS=7
run none PROF_FLAGS=0
run p4_none PROF_FLAGS=0 BSSM_B200_P=4 BSSM_B200_RANKS=64
run p2_none PROF_FLAGS=0 BSSM_B200_P=2 BSSM_B200_RANKS=128

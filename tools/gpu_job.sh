mkdir -p gpurun_out
DVP_MVS_LIB=$PWD/dvp_mvs_b200/libdvp_mvs_dupes.so python tools/count_dupes.py > gpurun_out/s17_dupes.log 2>&1; tail -4 gpurun_out/s17_dupes.log

mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_tex tools/ubench_tex.cu && /tmp/ubench_tex > gpurun_out/s21_ubench_tex.log 2>&1
tail -15 gpurun_out/s21_ubench_tex.log

# scratch GPU job (edited per call)
mkdir -p gpurun_out
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_tex tools/ubench_tex.cu && /tmp/ubench_tex > gpurun_out/s2_ubench_tex.log 2>&1
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
python bench.py --cpu-sample 0 > gpurun_out/s2_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s2_launches.csv python bench.py --steps 1 --warmup 1 --cpu-sample 0 > gpurun_out/s2_ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_gen_neighbours|k_depth_to_weak' -c 2 -o gpurun_out/s2_k4_k15 -f python bench.py --steps 1 --warmup 0 --cpu-sample 0 > gpurun_out/s2_ncu_full.log 2>&1
cat gpurun_out/s2_ubench_tex.log; tail -3 gpurun_out/s2_pytest.log; tail -1 gpurun_out/s2_bench.log

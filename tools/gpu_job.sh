# scratch GPU job (edited per call)
mkdir -p gpurun_out
for v in rw3 rw4 rw96; do
DVP_MVS_LIB=$PWD/dvp_mvs_b200/libdvp_mvs_$v.so python bench.py --cpu-sample 0 --steps 3 > gpurun_out/s4_bench_$v.log 2>&1
done
DVP_MVS_LIB=$PWD/dvp_mvs_b200/libdvp_mvs_rw4.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/s4_pytest_rw4.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s4_pytest_rw4.log
tail -3 gpurun_out/s4_pytest_rw4.log
for v in rw3 rw4 rw96; do tail -1 gpurun_out/s4_bench_$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$v', d['value'], d['per_stage_ms'])"; done

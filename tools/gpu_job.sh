# scratch GPU job (edited per call)
mkdir -p gpurun_out
python bench.py --cpu-sample 0 --steps 3 > gpurun_out/s9_bench_1gpu.log 2>&1; tail -1 gpurun_out/s9_bench_1gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['per_stage_ms'])"
python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/s9_pytest.log 2>&1; tail -2 gpurun_out/s9_pytest.log

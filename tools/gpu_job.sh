# scratch GPU job (edited per call)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s10_pytest.log 2>&1; tail -12 gpurun_out/s10_pytest.log | cut -c1-300
python bench.py --cpu-sample 0 --steps 3 > gpurun_out/s10_bench_1gpu.log 2>&1; tail -1 gpurun_out/s10_bench_1gpu.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['gpu_launches'], d['per_stage_ms'])"

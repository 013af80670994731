mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "overlapped" > gpurun_out/s18_pytest.log 2>&1; tail -12 gpurun_out/s18_pytest.log | cut -c1-300
python bench.py --cpu-sample 0 > gpurun_out/s18_bench.log 2>&1; tail -1 gpurun_out/s18_bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"

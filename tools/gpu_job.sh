# scratch GPU job (edited per call)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "many_source" > gpurun_out/s15_pytest.log 2>&1; tail -25 gpurun_out/s15_pytest.log | cut -c1-300

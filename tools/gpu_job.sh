# scratch GPU job (edited per call)
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s > gpurun_out/s5_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s5_pytest.log
tail -30 gpurun_out/s5_pytest.log | cut -c1-400

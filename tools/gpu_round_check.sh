set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/s2_pytest.log
python __graft_entry__.py smoke > gpurun_out/s2_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/s2_smoke.log
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/s2_bench_ref.log 2>&1
python bench.py > gpurun_out/s2_bench.log 2>&1
tail -3 gpurun_out/s2_pytest.log; tail -3 gpurun_out/s2_smoke.log; tail -1 gpurun_out/s2_bench_ref.log; tail -1 gpurun_out/s2_bench.log

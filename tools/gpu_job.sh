# scratch GPU job (edited per call)
mkdir -p gpurun_out
python -m pytest tests/test_scene.py -m gpu -q > gpurun_out/s16_pytest.log 2>&1; tail -3 gpurun_out/s16_pytest.log | cut -c1-300
python tools/farm_scene_demo.py --full 3111x2073 --views 6 --levels 3 > gpurun_out/s16_farm1.log 2>&1; tail -1 gpurun_out/s16_farm1.log | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 tools/farm_scene_demo.py --full 3111x2073 --views 6 --levels 3 > gpurun_out/s16_farm2.log 2>&1; tail -1 gpurun_out/s16_farm2.log | cut -c1-600

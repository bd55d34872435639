# round 2, call P (1 GPU, the last seconds of the budget): the k-epsilon step on the device
mkdir -p gpurun_out
timeout 40 python -m pytest tests/test_gpu_kepsilon.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02p_tests.log
cat gpurun_out/r02p_tests.log

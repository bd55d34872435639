# round 2, call N (1 GPU): final state of the face-sum kernels -- parity on the device, the kernel table
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02n_tests.log
timeout 400 python bench_kernels.py > gpurun_out/r02n_kernels.json 2> gpurun_out/r02n_kernels.txt
cat gpurun_out/r02n_tests.log; cat gpurun_out/r02n_kernels.txt

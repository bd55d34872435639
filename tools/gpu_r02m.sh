# round 2, call M (1 GPU): the face-sum / fvMatrix / MULES kernels with hoisted load scheduling -- parity on the device, then the table
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 > gpurun_out/r02m_tests.log
timeout 400 python bench_kernels.py > gpurun_out/r02m_kernels.json 2> gpurun_out/r02m_kernels.txt
cat gpurun_out/r02m_tests.log; cat gpurun_out/r02m_kernels.txt

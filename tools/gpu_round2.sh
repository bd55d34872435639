# 2-GPU validation: multi-rank tests + fused scalar tail on/off at two problem sizes
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py tests/test_gpu_pcg_paths.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r01g_tests.log
B200LDU_TAIL=1 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pcg_paths.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r01g_tests_tail.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$TR bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r01g_bench_n2.json 2>/dev/null
B200LDU_TAIL=1 $TR bench.py --gpus 2 --no-cpu-baseline > gpurun_out/r01g_bench_n2_tail.json 2>/dev/null
$TR bench.py --gpus 2 --n 160 --no-cpu-baseline > gpurun_out/r01g_bench_n2_160.json 2>/dev/null
B200LDU_TAIL=1 $TR bench.py --gpus 2 --n 160 --no-cpu-baseline > gpurun_out/r01g_bench_n2_160_tail.json 2>/dev/null
B200LDU_TAIL=1 python bench.py --no-cpu-baseline > gpurun_out/r01g_bench_n1_tail.json 2>/dev/null
B200LDU_TAIL=1 python bench.py --n 128 --no-cpu-baseline > gpurun_out/r01g_bench_n128_tail.json 2>/dev/null
cat gpurun_out/r01g_tests.log gpurun_out/r01g_tests_tail.log
for f in gpurun_out/r01g_bench_*.json; do echo $f; cut -c1-200 $f; done

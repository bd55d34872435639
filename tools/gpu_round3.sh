# 8-GPU: scaling point + multi-rank tests at 8 ranks
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r01h_bench_n8.json 2> gpurun_out/r01h_bench_n8.err
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r01h_tests.log
timeout 200 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --no-cpu-baseline > gpurun_out/r01h_bench_n4.json 2> gpurun_out/r01h_bench_n4.err
cat gpurun_out/r01h_tests.log
for f in gpurun_out/r01h_bench_*.json; do echo $f; cut -c1-220 $f; done
tail -3 gpurun_out/r01h_bench_n8.err

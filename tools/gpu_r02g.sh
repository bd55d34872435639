# round 2, call G (8 GPUs): multi-GPU tests at 2/4/8 ranks, strong scaling at 4 and 8 with both scalar-step structures
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_zzz_fvm_gpu.py -m gpu -q -k "multi" 2>&1 | tail -12 > gpurun_out/r02g_tests_multi.log
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r02g_bench_n8.json 2> gpurun_out/r02g_bench_n8.err
B200LDU_PCG_DEFERRED=0 timeout 300 $TR --nproc-per-node 8 --master-port 29522 bench.py --gpus 8 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/r02g_bench_n8_4launch.json 2> gpurun_out/r02g_bench_n8_4launch.err
timeout 300 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --no-cpu-baseline --no-secondary > gpurun_out/r02g_bench_n4.json 2> gpurun_out/r02g_bench_n4.err
B200LDU_PCG_DEFERRED=0 timeout 300 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --no-cpu-baseline --no-secondary --no-parity > gpurun_out/r02g_bench_n4_4launch.json 2> gpurun_out/r02g_bench_n4_4launch.err
cat gpurun_out/r02g_tests_multi.log
for f in gpurun_out/r02g_bench_*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:d.get(k) for k in ("value","value_solver_only","parity","comm","gpu_launches")}, d["e2e"]["value"], d.get("secondary"))
except Exception as e: print("ERR", sys.argv[1], e)
PY
done
tail -3 gpurun_out/r02g_bench_n8.err

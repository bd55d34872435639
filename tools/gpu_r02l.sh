# round 2, call L (8 GPUs): the final code at 8 ranks -- strong-scaling bench (parity gate, secondary legs), then the multi-GPU tests
# (MULES over processor patches new) and the CMULES device tests
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 150 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --no-cpu-baseline > gpurun_out/r02l_bench_n8.json 2> gpurun_out/r02l_bench_n8.err
timeout 200 python -m pytest "tests/test_gpu_mules.py::test_multi_gpu_mules[8]" "tests/test_gpu_multi.py::test_multi_gpu_solvers[8]" "tests/test_zzz_fvm_gpu.py::test_multi_gpu_icofoam[8]" "tests/test_gpu_mules.py::test_multi_gpu_mules[2]" tests/test_gpu_mules.py::test_cmules_on_the_device -m gpu -q --durations=8 2>&1 | tail -22 > gpurun_out/r02l_tests_multi.log
cat gpurun_out/r02l_tests_multi.log
python - gpurun_out/r02l_bench_n8.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:d.get(k) for k in ("value","value_solver_only","parity","comm","layout","gpu_launches")}, d["e2e"]["value"], d.get("secondary"))
except Exception as e: print("ERR", sys.argv[1], e)
PY
tail -3 gpurun_out/r02l_bench_n8.err

# round 2, call F (2 GPUs): gate-word deferred steps, cyclic GAMG, limiters, lduops, batched face sums
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r02f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > gpurun_out/r02f_smoke.log 2>&1
timeout 300 python tools/bench_engine.py > gpurun_out/r02f_engine.jsonl 2> gpurun_out/r02f_engine.err
timeout 600 python bench_kernels.py > gpurun_out/r02f_kernels.json 2> gpurun_out/r02f_kernels.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --no-cpu-baseline --no-secondary > gpurun_out/r02f_bench_n2.json 2> gpurun_out/r02f_bench_n2.err
tail -12 gpurun_out/r02f_tests.log; tail -2 gpurun_out/r02f_smoke.log; cat gpurun_out/r02f_engine.jsonl; tail -3 gpurun_out/r02f_engine.err
head -24 gpurun_out/r02f_kernels.txt
python - gpurun_out/r02f_bench_n2.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:d.get(k) for k in ("value","value_solver_only","parity","comm","gpu_launches")}, d["e2e"]["value"])
except Exception as e: print("ERR", e)
PY
tail -5 gpurun_out/r02f_bench_n2.err

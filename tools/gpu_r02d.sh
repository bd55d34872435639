# round 2, call D (2 GPUs): two-launch fused PCG (prologue scalar step + peer all-reduce) + lattice tile embedding
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pcg_paths.py tests/test_gpu_multi.py tests/test_zzz_fvm_gpu.py tests/test_reference_binding.py tests/test_gpu_gamg.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02d_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > gpurun_out/r02d_smoke.log 2>&1
: > gpurun_out/r02d_engine_ab.jsonl
for v in "" _g6m4 _g4m6 _g2m6; do
  B200LDU_LIB=rapidcfd-dev_b200/lib/libb200ldu$v.so timeout 300 python tools/bench_engine.py >> gpurun_out/r02d_engine_ab.jsonl 2>> gpurun_out/r02d_engine_ab.err
done
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --no-cpu-baseline --no-secondary > gpurun_out/r02d_bench_n2.json 2> gpurun_out/r02d_bench_n2.err
timeout 600 $TR bench.py --gpus 2 --n 160 --no-cpu-baseline --no-secondary > gpurun_out/r02d_bench_n2_160.json 2>> gpurun_out/r02d_bench_n2.err
tail -6 gpurun_out/r02d_tests.log; tail -2 gpurun_out/r02d_smoke.log; cat gpurun_out/r02d_engine_ab.jsonl; tail -3 gpurun_out/r02d_engine_ab.err
for f in gpurun_out/r02d_bench_n2*.json; do python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], {k:d.get(k) for k in ("value","value_solver_only","parity","comm")}, d["e2e"]["value"])
except Exception as e: print("ERR", e)
PY
done
tail -5 gpurun_out/r02d_bench_n2.err

# round 2, call B (1 GPU): first run of the compressed-column engine
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02b_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > gpurun_out/r02b_smoke.log 2>&1
python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err
B200LDU_LIB=rapidcfd-dev_b200/lib/libb200ldu_minb6.so python bench.py --no-cpu-baseline --no-secondary > gpurun_out/r02b_bench_n1_minb6.json 2> gpurun_out/r02b_bench_n1_minb6.err
python bench.py --n 128 --no-cpu-baseline --no-secondary > gpurun_out/r02b_bench_n128.json 2>/dev/null
B200LDU_LIB=rapidcfd-dev_b200/lib/libb200ldu_minb6.so python bench.py --n 128 --no-cpu-baseline --no-secondary > gpurun_out/r02b_bench_n128_minb6.json 2>/dev/null
python bench_kernels.py > gpurun_out/r02b_kernels.json 2> gpurun_out/r02b_kernels.txt
python tools/diag_hist.py > gpurun_out/r02b_diag_hist.txt 2>&1
tail -5 gpurun_out/r02b_tests.log; tail -2 gpurun_out/r02b_smoke.log
for f in gpurun_out/r02b_bench_*.json; do echo $f; python - $f <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print({k:d.get(k) for k in ("value","value_solver_only","parity")}, d["roofline"]["amul"]["ms_per_launch"], d["roofline"]["amul"]["frac"], d["e2e"]["value"], d.get("secondary"))
except Exception as e: print("ERR", e)
PY
done
tail -3 gpurun_out/r02b_bench_n1.err

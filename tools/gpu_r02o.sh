# round 2, call O (1 GPU): the final tree -- GPU tests and smoke()
set -x
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -q -x 2>&1 | tail -6 > gpurun_out/r02o_tests.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > gpurun_out/r02o_smoke.log 2>&1
cat gpurun_out/r02o_tests.log; tail -2 gpurun_out/r02o_smoke.log

# round 2, call C (1 GPU): persistent PCG first run + engine A/B (group size, CTAs/SM, column mode) + ncu of Amul
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_pcg_paths.py tests/test_zzz_fvm_gpu.py -m gpu -q -x 2>&1 | tail -15 > gpurun_out/r02c_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > gpurun_out/r02c_smoke.log 2>&1
: > gpurun_out/r02c_engine_ab.jsonl
for v in "" _g8m4 _g6m4 _g6m5 _c1g4m6 _c1g8m4; do
  B200LDU_LIB=rapidcfd-dev_b200/lib/libb200ldu$v.so timeout 300 python tools/bench_engine.py >> gpurun_out/r02c_engine_ab.jsonl 2>> gpurun_out/r02c_engine_ab.err
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"AmulOp<0>|pcg_persistent" -s 6 -c 3 -f -o gpurun_out/r02c_amul python tools/bench_engine.py --sizes 256 --reps 2 > gpurun_out/r02c_ncu.log 2>&1
tail -6 gpurun_out/r02c_tests.log; tail -2 gpurun_out/r02c_smoke.log; cat gpurun_out/r02c_engine_ab.jsonl; tail -3 gpurun_out/r02c_engine_ab.err

# round 2, call K (1 GPU): kernel table with the MULES rows, ncu --set full of the face-sum / fvMatrix / GAMG / Jacobi / MULES kernels
# (reduced to a per-launch summary; the .ncu-rep stays in /tmp), compute-sanitizer memcheck of the round's new kernels
set -x
mkdir -p gpurun_out
timeout 500 python bench_kernels.py > gpurun_out/r02k_kernels.json 2> gpurun_out/r02k_kernels.txt
timeout 600 ncu --set full --clock-control none -k regex:"JacobiOp|ResidualOp|CoeffSumOp|surface_integrate|gauss_grad|interpolate_linear|laplacian_upper|convection_faces|neg_sum_diag|grad_linear|flux_linear|faceH_kernel|H_kernel|relax_kernel|A_kernel|flux_internal|boundary_|residual_source|restrict|prolong|agg_|dense_apply|row_sum|limiter|limited_weights|mules_" -c 70 -f -o /tmp/r02k_kernels python bench_kernels.py --n 128 --reps 1 > gpurun_out/r02k_ncu_kernels.log 2>&1
ncu -i /tmp/r02k_kernels.ncu-rep --page raw --csv > /tmp/r02k_raw.csv 2>/dev/null
python tools/ncu_summary.py /tmp/r02k_raw.csv > gpurun_out/r02k_ncu_kernels_summary.csv
timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_mules.py tests/test_gpu_limiters.py tests/test_lduops.py -m gpu -q -x -k "not multi_gpu" > gpurun_out/r02k_sanitizer_memcheck.txt 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r02k_sanitizer_memcheck.txt
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE-OK')" > gpurun_out/r02k_sanitizer_smoke.txt 2>&1
echo "sanitizer rc=$?" >> gpurun_out/r02k_sanitizer_smoke.txt
grep -i "mules" gpurun_out/r02k_kernels.txt
tail -4 gpurun_out/r02k_sanitizer_memcheck.txt; tail -4 gpurun_out/r02k_sanitizer_smoke.txt
wc -l gpurun_out/r02k_ncu_kernels_summary.csv; du -sh gpurun_out

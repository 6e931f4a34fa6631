set -x
T="tests/test_gpu_solver.py::test_linearize_matches_oracle[200-anchor-True-2] tests/test_gpu_solver.py::test_linearize_matches_oracle[200-anchor-True-1] tests/test_gpu_solver.py::test_solve_matches_oracle[200-anchor-2] tests/test_gpu_marg.py::test_margin_old_matches_oracle[cholesky-anchor-200] tests/test_gpu_solver.py::test_device_preintegration_matches_oracle tests/test_gpu_solver.py::test_config4_wheel_and_lidar_planes_match_oracle[200-777-False-2]"
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest $T -q -m gpu -x > gpurun_out/compute_sanitizer_memcheck_r2.log 2>&1
tail -5 gpurun_out/compute_sanitizer_memcheck_r2.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python -m pytest "tests/test_gpu_solver.py::test_linearize_matches_oracle[200-anchor-True-2]" "tests/test_gpu_marg.py::test_margin_old_matches_oracle[cholesky-anchor-200]" tests/test_gpu_solver.py::test_device_preintegration_matches_oracle -q -m gpu -x > gpurun_out/compute_sanitizer_racecheck_r2.log 2>&1
tail -8 gpurun_out/compute_sanitizer_racecheck_r2.log

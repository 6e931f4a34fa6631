set -x
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:k_linearize|k_solve2|k_backsub|k_cand_eval|k_nonvis|k_imu_preintegrate" -s 11 -c 6 -f -o gpurun_out/ncu_solver_r2b python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-marginalize --no-lk --no-config4 > gpurun_out/ncu_solver_r2b.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_linearize_ws|k_step|k_marg_eig|k_marg_build" -s 6 -c 4 -f -o gpurun_out/ncu_window_r2b python scripts/dbg/prof_window.py > gpurun_out/ncu_window_r2b.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 110 -c 170 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-marginalize --no-lk --no-config4 > gpurun_out/launches_r2b.log 2>&1
ls -la gpurun_out | tail -8

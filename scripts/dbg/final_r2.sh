set -x
timeout 1400 python bench.py > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; tail -c 200 gpurun_out/bench_r2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_r2.json 2>/dev/null
timeout 500 ncu --set full --clock-control none --import-source on -k "regex:k_linearize|k_solve2|k_backsub|k_cand_eval|k_nonvis|k_imu_preintegrate" -s 10 -c 7 -f -o gpurun_out/ncu_solver_r2c python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-marginalize --no-lk --no-config4 > gpurun_out/ncu_solver_r2c.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_linearize_ws|k_step|k_marg_eig|k_marg_build" -s 6 -c 4 -f -o gpurun_out/ncu_window_r2c python scripts/dbg/prof_window.py > gpurun_out/ncu_window_r2c.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s 110 -c 170 --csv --log-file gpurun_out/launches_r2c.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-marginalize --no-lk --no-config4 > gpurun_out/launches_r2c.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
ls -la gpurun_out | tail -6

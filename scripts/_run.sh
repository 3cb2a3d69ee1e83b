# Round-2 closing measurements on one B200 (run under gpurun from the repository root); the
# artefacts land in gpurun_out/ and are copied to profiles/ by hand.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; tail -3 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python bench.py > gpurun_out/r2c_bench_n1.json 2> gpurun_out/bench.err || tail -5 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2c_bench_reference_arm.json 2>> gpurun_out/bench.err
# launch list of the bench command (per-launch times are cold-cache and serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2c_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log | cut -c1-300
# DRAM traffic and pipe figures of one launch of the bench command itself
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum --clock-control none -k regex:track_kernel -s 8 -c 1 --csv --log-file gpurun_out/r2c_ncu_dram_bench_launch.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_d.log 2>&1
tail -4 gpurun_out/r2c_ncu_dram_bench_launch.csv | cut -c1-400
# source-level captures: C2 (one wave of the default kernel, work-queue path) and C5
ncu --set full --clock-control none --import-source on -k regex:track_kernel -c 1 -f -o gpurun_out/r2c_c2 python scripts/profile_target.py 227328 20 4 c2 > gpurun_out/ncu_c2.log 2>&1; tail -1 gpurun_out/ncu_c2.log
ncu --set full --clock-control none --import-source on -k regex:track_kernel -c 1 -f -o gpurun_out/r2c_c5 python scripts/profile_target.py 151552 40 2 c5 > gpurun_out/ncu_c5.log 2>&1; tail -1 gpurun_out/ncu_c5.log
python scripts/sweep_n.py gpurun_out/r2c_sweep_n.json 30 > gpurun_out/sweep.log 2>&1; tail -2 gpurun_out/sweep.log
python tests/accuracy_study.py r2c > gpurun_out/acc.log 2>&1; tail -3 gpurun_out/acc.log

mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/b2.json 2> gpurun_out/b2.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_ncu_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_l.log 2>&1
tail -2 gpurun_out/ncu_l.log | cut -c1-300
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:track_kernel -s 8 -c 1 --csv --log-file gpurun_out/r2_ncu_dram_bench_launch.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_d.log 2>&1
tail -3 gpurun_out/r2_ncu_dram_bench_launch.csv | cut -c1-400
python scripts/sweep_n.py gpurun_out/r2_sweep_n.json 30 > gpurun_out/sweep.log 2>&1; tail -3 gpurun_out/sweep.log
python tests/accuracy_study.py r2 > gpurun_out/acc.log 2>&1; tail -5 gpurun_out/acc.log

# Closing measurements of round 2 after the per-chunk barrier of the beam-field kernels (r2d; run
# under gpurun from the repository root); artefacts land in gpurun_out/ and are copied to profiles/
# by hand.  The thin-lens kernels (C2, C4) are the ones of the r2c run: their ncu launch list, DRAM
# figures, source-level capture, N sweep and accuracy study (profiles/r2c_*) were not repeated;
# the full r2c command list is in the history of this file.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; tail -3 gpurun_out/pytest_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log | cut -c1-200
python bench.py > gpurun_out/r2d_bench_n1.json 2> gpurun_out/bench.err || tail -5 gpurun_out/bench.err
cut -c1-300 gpurun_out/r2d_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2d_bench_reference_arm.json 2>> gpurun_out/bench.err
cut -c1-300 gpurun_out/r2d_bench_reference_arm.json
# C5 at full size (1e6 particles x 1e4 turns, BeamMonitor), clocks sampled alongside
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_throttle_reasons.active --format=csv -lms 500 > gpurun_out/full_clocks.csv 2>/dev/null &
SMI=$!
python scripts/run_c5_psb.py > gpurun_out/c5_full.log 2>&1; tail -2 gpurun_out/c5_full.log | cut -c1-300
kill $SMI

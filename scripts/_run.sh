mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; tail -5 gpurun_out/r2_gputests.log
python scripts/probe_variants.py --configs c2,c5,c3 base defer > gpurun_out/r2_probe4.txt 2>&1
cat gpurun_out/r2_probe4.txt | cut -c1-160

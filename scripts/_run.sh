mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests.log 2>&1; tail -5 gpurun_out/r2_gputests.log
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n1.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],'strict',d['strict']['value'])
for k,v in d['configs'].items(): print(k, v['value'], v['roofline']['frac'], v['kernel_shape'])
PY

# quick GPU check after a kernel change: the step-level parity tests, then the bench line without aux / CPU legs
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sage.py tests/test_gpu_parity_reddit.py tests/test_gpu_fullsize.py -x -q > gpurun_out/qc_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/qc_tests.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-aux --no-parity ${BENCH_EXTRA:-} > gpurun_out/qc_bench.json 2> gpurun_out/qc_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/qc_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/qc_bench.json').read().strip().splitlines()[-1])
print(d['dtype'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']))
for a in d['alt'] or []: print(a['dtype'], round(a['value']), a['ms_per_step'])
print({k:(v['ms'], v.get('gbs') or v.get('tflops')) for k,v in d['stages'].items()})
PY

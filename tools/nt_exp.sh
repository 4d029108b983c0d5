set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q > gpurun_out/el_t1.log 2>&1; echo "gemm tests rc=$?"; tail -3 gpurun_out/el_t1.log
: > gpurun_out/nt_exp3.log
for d in 3 0; do OGL_GEMM_DBG=$d timeout 300 python tools/nt_exp2.py 2>&1 | grep "cg=2" >> gpurun_out/nt_exp3.log; done
for d in 0 1 2; do OGL_GEMM_DBG=$d timeout 200 python tools/nt_exp.py >> gpurun_out/nt_exp3.log 2>&1; done
OGL_GEMM_LOADER=0 TAG=one_producer timeout 200 python tools/nt_exp.py >> gpurun_out/nt_exp3.log 2>&1
cat gpurun_out/nt_exp3.log
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-aux --no-parity > gpurun_out/el_b1.json 2> gpurun_out/el_b1.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/el_b1.json').read().strip().splitlines()[-1])
print(d['dtype'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'])
for a in d['alt'] or []: print(a['dtype'], a['value'], a['ms_per_step'])
for k,v in d['stages'].items(): print(k, v)
PY

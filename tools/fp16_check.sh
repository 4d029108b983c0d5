# focused GPU check of the fp16 mode: kernel tests, step parity at the bench shape, then the bench line (all three arithmetic modes)
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_sage.py tests/test_gpu_parity_reddit.py -x -q -s -k "fp16" > gpurun_out/fp16_t1.log 2>&1
echo "tests rc=$?"; grep -E "passed|failed|Error|error" gpurun_out/fp16_t1.log | tail -5; grep "PARITY fp16" gpurun_out/fp16_t1.log | cut -c1-600
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-aux > gpurun_out/fp16_b1.json 2> gpurun_out/fp16_b1.err
echo "bench rc=$?"; tail -c 400 gpurun_out/fp16_b1.err; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/fp16_b1.json').read().strip().splitlines()[-1])
    print(d['dtype'], d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'parity', d['parity']['meets_rtol_1e-3'], d['parity']['logits_max_err_of_scale'], d['parity']['grad_rel_fro_pinned_max'])
    print('trained', d['parity_trained']['meets_rtol_1e-3'], d['parity_trained']['logits_max_err_of_scale'], d['parity_trained']['grad_rel_fro_pinned_max'])
    for a in d['alt'] or []:
        print(a['dtype'], a['value'], a['ms_per_step'], a['parity']['logits_max_err_of_scale'], a['parity']['grad_rel_fro_pinned_max'])
    print({k:v for k,v in d['roofline'].items() if k in ('launch','ms_per_launch','achieved','frac','frac_burst')})
except Exception as e:
    print('parse failed', e)
PY

timeout 300 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_sage.py -x -q > gpurun_out/t21.log 2>&1; tail -4 gpurun_out/t21.log
timeout 120 python tools/tc_exp.py 2>&1 | grep "^TN"
timeout 200 python bench.py --no-cpu-baseline --no-aux > gpurun_out/b21.json 2> gpurun_out/b21.err
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/b21.json").read().strip().splitlines()[-1]); print("b21", round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), {k:(v["ms"], v.get("tflops")) for k,v in d["stages"].items() if "dW" in k})
except Exception as e: print("ERR", e)
PY
tail -2 gpurun_out/b21.err

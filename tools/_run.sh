timeout 600 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_sage.py tests/test_gpu_peer.py -x -q > gpurun_out/t19.log 2>&1; tail -5 gpurun_out/t19.log
python bench.py --no-cpu-baseline --no-aux > gpurun_out/b19.json 2> gpurun_out/b19.err
python - <<'PY'
import json
for f in ("b19",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), {k:(v["ms"], v.get("tflops")) for k,v in d["stages"].items() if "dW" in k or "gemm" in k})
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/b19.err

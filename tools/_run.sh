python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_sage.py -x -q > gpurun_out/t18.log 2>&1; tail -3 gpurun_out/t18.log
for v in 0 1 -1; do OGL_PRE_PRIO=$v python bench.py --no-cpu-baseline --no-aux > gpurun_out/b18_p$v.json 2> gpurun_out/b18_p$v.err; done
python bench.py --no-cpu-baseline --no-aux --no-pipeline > gpurun_out/b18_np.json 2> gpurun_out/b18_np.err
python - <<'PY'
import json
for f in ("b18_p0","b18_p1","b18_p-1","b18_np"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), {k:v["ms"] for k,v in list(d["stages"].items())[:6]})
    except Exception as e: print(f, "ERR", e)
PY

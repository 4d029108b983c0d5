timeout 600 python -m pytest tests/test_gpu_sage.py tests/test_gpu_peer.py tests/test_gpu_trainers.py -x -q > gpurun_out/t20.log 2>&1; tail -4 gpurun_out/t20.log
python bench.py --no-cpu-baseline --no-aux > gpurun_out/b20.json 2> gpurun_out/b20.err
python - <<'PY'
import json
for f in ("b20",):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1]); print(f, round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["host_enqueue_ms_per_step"])
    except Exception as e: print(f, "ERR", e)
PY
tail -3 gpurun_out/b20.err

set -u
mkdir -p gpurun_out
L=online-gnn-learning_b200/libogl_b200.so
cp $L /tmp/lib_s5.so
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q > gpurun_out/ep_t1.log 2>&1; echo "gemm tests (default) rc=$?"; tail -2 gpurun_out/ep_t1.log
OGL_NT_ORDER=1 timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q > gpurun_out/ep_t2.log 2>&1; echo "gemm tests (order 1) rc=$?"; tail -2 gpurun_out/ep_t2.log
: > gpurun_out/nt_exp4.log
for d in 0 1 2; do OGL_GEMM_DBG=$d TAG=s5 timeout 200 python tools/nt_exp.py >> gpurun_out/nt_exp4.log 2>&1; done
OGL_NT_ORDER=1 TAG=s5_order1 timeout 200 python tools/nt_exp.py >> gpurun_out/nt_exp4.log 2>&1
for pf in 4 8 16; do OGL_GEMM_PF_NT=$pf TAG=s5_pf$pf timeout 200 python tools/nt_exp.py >> gpurun_out/nt_exp4.log 2>&1; done
cp tools/exp/libogl_b200_s6.so $L
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -x -q > gpurun_out/ep_t3.log 2>&1; echo "gemm tests (6 stages) rc=$?"; tail -2 gpurun_out/ep_t3.log
for d in 0 1; do OGL_GEMM_DBG=$d TAG=s6 timeout 200 python tools/nt_exp.py >> gpurun_out/nt_exp4.log 2>&1; done
OGL_NT_ORDER=1 TAG=s6_order1 timeout 200 python tools/nt_exp.py >> gpurun_out/nt_exp4.log 2>&1
OGL_GEMM_PF_NT=8 TAG=s6_pf8 timeout 200 python tools/nt_exp.py >> gpurun_out/nt_exp4.log 2>&1
grep -v "n=256\|m=18432" gpurun_out/nt_exp4.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-aux --no-parity --no-alt"
timeout 300 $B > gpurun_out/ep_b_s6.json 2> gpurun_out/ep_b_s6.err; echo "bench s6 rc=$?"
OGL_NT_ORDER=1 timeout 300 $B > gpurun_out/ep_b_s6o1.json 2> gpurun_out/ep_b_s6o1.err
cp /tmp/lib_s5.so $L
timeout 300 $B > gpurun_out/ep_b_s5.json 2> gpurun_out/ep_b_s5.err; echo "bench s5 rc=$?"
OGL_NT_ORDER=1 timeout 300 $B > gpurun_out/ep_b_s5o1.json 2> gpurun_out/ep_b_s5o1.err
python - <<'PY'
import json
for n in ('s5','s5o1','s6','s6o1'):
    try:
        d=json.loads(open('gpurun_out/ep_b_%s.json'%n).read().strip().splitlines()[-1])
        st=d['stages']
        print(n, d['dtype'], round(d['value']), d['ms_per_step'], 'e2e', round(d['e2e']['value']), {k:st[k]['ms'] for k in ('l0.pool_gemm','l0.out_gemm','l1.dx_gemm','l0.dneigh_gemm','l1.pool_gemm','l0.dW_pool','l0.dW_group')})
    except Exception as e: print(n,'failed',e)
PY

# Round-end validation on one B200 (profiles/README.md): smoke, default bench (both arms), ncu launch list and full captures.
# usage: bash tools/final_run.sh [quick]      (the GPU test suite is a separate call: python -m pytest tests -m gpu)
set -u
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.json; echo
if [ "${1:-}" != "quick" ]; then
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; cut -c1-200 gpurun_out/final_ref.json
fi
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-aux --no-parity --no-alt"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r2b.csv $B > gpurun_out/ncu_l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_tn_tc -s 10 -c 2 -f -o gpurun_out/prof_tn_r2b $B > gpurun_out/ncu_tn.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_gemm_nt_tc -s 21 -c 2 -f -o gpurun_out/prof_nt_r2b $B > gpurun_out/ncu_nt.log 2>&1
if [ "${1:-}" != "quick" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:k_pool_bwd|k_segmax_fwd|k_gather_rows" -s 6 -c 5 -f -o gpurun_out/prof_mem_r2b $B > gpurun_out/ncu_mem.log 2>&1
fi
ls -la gpurun_out/*r2b*

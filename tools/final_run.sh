# Round-end validation on one B200: GPU test suite, smoke, default bench (both arms), ncu launch list and full captures (profiles/README.md)
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 400 gpurun_out/final_bench.json; echo
if [ "${1:-}" != "quick" ]; then
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err; cut -c1-200 gpurun_out/final_ref.json
fi
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-aux"
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r1d.csv $B > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tn_tc -s 10 -c 2 -f -o gpurun_out/prof_tn_r1d $B > gpurun_out/ncu_tn.log 2>&1
if [ "${1:-}" != "quick" ]; then
ncu --set full --clock-control none --import-source on -k regex:k_gemm_nt_tc -s 21 -c 2 -f -o gpurun_out/prof_nt_r1d $B > gpurun_out/ncu_nt.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_pool_bwd|k_segmax_fwd|k_gather_rows|k_sample" -s 21 -c 7 -f -o gpurun_out/prof_mem_r1d $B > gpurun_out/ncu_mem.log 2>&1
fi
ls -la gpurun_out/*r1d*

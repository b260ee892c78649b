set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/h_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --layers gpurun_out/h_layers_h5.txt > gpurun_out/h_bench_h5.log 2>&1; echo "h5 rc=$?"
timeout 200 python bench.py --workload h7 --no-cpu-baseline --layers gpurun_out/h_layers_h7.txt > gpurun_out/h_bench_h7.log 2>&1; echo "h7 rc=$?"
timeout 200 python bench.py --workload m7 --no-cpu-baseline --layers gpurun_out/h_layers_m7.txt > gpurun_out/h_bench_m7.log 2>&1; echo "m7 rc=$?"
timeout 200 python bench.py --workload m9 --no-cpu-baseline --layers gpurun_out/h_layers_m9.txt > gpurun_out/h_bench_m9.log 2>&1; echo "m9 rc=$?"
timeout 200 python bench.py --workload fill --no-cpu-baseline > gpurun_out/h_bench_fill.log 2>&1; echo "fill rc=$?"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-modes > gpurun_out/h_ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 100 python scripts/tc5s_waits.py > gpurun_out/h_tc5s_waits.txt 2>&1
timeout 100 python scripts/wgrad_line_waits.py > gpurun_out/h_wl_waits.txt 2>&1
for w in h5 h7 m7 m9 fill; do tail -1 gpurun_out/h_bench_$w.log | cut -c1-200; done

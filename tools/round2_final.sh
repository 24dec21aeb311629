python -m pytest tests -m gpu -q -s > gpurun_out/r2u_gpu_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2u_gpu_tests.log; grep -E "passed|failed" gpurun_out/r2u_gpu_tests.log | tail -2
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1; tail -1 gpurun_out/r2u_smoke.log
python bench.py > gpurun_out/r2u_bench_n1.json 2> gpurun_out/r2u_bench_n1.err; tail -c 200 gpurun_out/r2u_bench_n1.err
python bench.py --optimiser lbfgs --no-cpu-baseline > gpurun_out/r2u_bench_n1_lbfgs.json 2> gpurun_out/r2u_bench_n1_lbfgs.err
python bench.py --workload cfg5 --steps 2 --warmup 3 --no-family-pass > gpurun_out/r2u_bench_cfg5.json 2> gpurun_out/r2u_bench_cfg5.err
python bench.py --workload realmask --steps 2 --warmup 3 --no-family-pass > gpurun_out/r2u_bench_realmask.json 2> gpurun_out/r2u_bench_realmask.err

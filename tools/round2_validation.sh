python -m pytest tests -m gpu -q > gpurun_out/r2l_gpu_tests.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2l_gpu_tests.log; tail -4 gpurun_out/r2l_gpu_tests.log
python bench.py > gpurun_out/r2l_bench_n1.json 2> gpurun_out/r2l_bench_n1.err; tail -c 300 gpurun_out/r2l_bench_n1.err
python bench.py --steps 1 --warmup 1 --no-full-day --no-family-pass --no-cpu-baseline > gpurun_out/r2l_plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 3000 --csv --log-file gpurun_out/r2l_launches.csv python bench.py --steps 1 --warmup 1 --no-full-day --no-family-pass --no-cpu-baseline > gpurun_out/r2l_ncu_bench.log 2>&1
export OI_GROUPS=1
python tools/eval_bench.py 16 2 > gpurun_out/r2l_eval.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2l_traffic.csv python tools/eval_bench.py 16 1 > gpurun_out/r2l_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_chol_update -s 10 -c 2 -o gpurun_out/r2l_prof_chol python tools/eval_bench.py 16 1 > gpurun_out/r2l_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_lauum_trace -s 10 -c 2 -o gpurun_out/r2l_prof_lauum python tools/eval_bench.py 16 1 > gpurun_out/r2l_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_trtri -s 10 -c 2 -o gpurun_out/r2l_prof_trtri python tools/eval_bench.py 16 1 > gpurun_out/r2l_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_chol_panel -s 10 -c 2 -o gpurun_out/r2l_prof_panel python tools/eval_bench.py 16 1 > gpurun_out/r2l_ncu5.log 2>&1
cat gpurun_out/r2l_eval.log

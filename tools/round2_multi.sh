# 8-GPU measurements (one box): the whole day sharded by LPT (bench step at N=8), the season sharded by day
P=29500
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 3 --warmup 1 > gpurun_out/r2m_bench_n8.json 2> gpurun_out/r2m_bench_n8.err; tail -c 300 gpurun_out/r2m_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((P+1)) bench.py --gpus 8 --steps 3 --warmup 1 --optimiser lbfgs > gpurun_out/r2m_bench_n8_lbfgs.json 2> gpurun_out/r2m_bench_n8_lbfgs.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((P+2)) tools/season_bench.py --days 16 --optimiser lbfgs 2>/dev/null | tail -1 > gpurun_out/r2m_season_n8_lbfgs.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((P+3)) tools/season_bench.py --days 16 --cell-stride 8 2>/dev/null | tail -1 > gpurun_out/r2m_season_n8_cg_stride8.json
cat gpurun_out/r2m_season_n8_lbfgs.json gpurun_out/r2m_season_n8_cg_stride8.json

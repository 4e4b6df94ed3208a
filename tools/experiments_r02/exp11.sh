python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -3
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-sub-records > gpurun_out/r02q_bench_n2.json 2> gpurun_out/r02q_bench_n2.err
tail -2 gpurun_out/r02q_bench_n2.err

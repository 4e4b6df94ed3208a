python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02n_bench_n2.json 2> gpurun_out/r02n_bench_n2.err
tail -3 gpurun_out/r02n_bench_n2.err
python bench.py --steps 5 --warmup 3 > gpurun_out/r02n_bench_n1.json 2> gpurun_out/r02n_bench_n1.err
tail -3 gpurun_out/r02n_bench_n1.err

mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 4 --steps 20 --warmup 5 --profile > gpurun_out/s18_bench4.json 2> gpurun_out/s18_bench4.err; echo "rc $?" >> gpurun_out/s18_bench4.err

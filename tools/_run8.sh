mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29701 bench.py --gpus 8 --steps 20 --warmup 5 --profile > gpurun_out/s8_bench8_1.json 2> gpurun_out/s8_bench8_1.err; echo "rc $?" >> gpurun_out/s8_bench8_1.err
$T --master-port 29702 bench.py --gpus 8 --steps 200 --warmup 5 --profile --secondary '' > gpurun_out/s8_bench8_2.json 2> gpurun_out/s8_bench8_2.err; echo "rc $?" >> gpurun_out/s8_bench8_2.err
$T --master-port 29703 tools/check_sharded.py --workload S64_disc --steps 20 --plan replicated --packed > gpurun_out/s8_check_repl.txt 2>&1; echo "rc $?" >> gpurun_out/s8_check_repl.txt
$T --master-port 29704 tools/check_sharded.py --workload S32_disc --steps 30 --plan rows > gpurun_out/s8_check_rows.txt 2>&1; echo "rc $?" >> gpurun_out/s8_check_rows.txt

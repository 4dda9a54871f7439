mkdir -p gpurun_out
nvidia-smi -L | wc -l
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "rc=$?"; cut -c1-400 gpurun_out/bench_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_4gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --streams 4096 --groups 2 --steps 2000 --warmup 50 > gpurun_out/bench_8gpu_4096.json 2> gpurun_out/bench_8gpu_4096.err; echo "rc=$?"; cut -c1-300 gpurun_out/bench_8gpu_4096.json

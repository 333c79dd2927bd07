# gpurun --gpus N -- 'bash profiles/run_multi_gpu.sh N': the 2-GPU K5 tests and the bench line at N GPUs
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests/test_p2p_multi_gpu.py -x -q -m gpu > gpurun_out/s2_p2p_test_${N}gpu.log 2>&1; echo rc=$? >> gpurun_out/s2_p2p_test_${N}gpu.log; tail -2 gpurun_out/s2_p2p_test_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/s2_bench_${N}gpu.json 2> gpurun_out/s2_bench_${N}gpu.err; echo bench rc=$?
tail -c 600 gpurun_out/s2_bench_${N}gpu.json

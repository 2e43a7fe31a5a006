python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/n2_tests.txt; cat gpurun_out/n2_tests.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-newton --no-spot > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
tail -c 1500 gpurun_out/n2_bench.json
compute-sanitizer --tool memcheck python tools/sanitize_once.py 2>&1 | tail -8 > gpurun_out/n2_memcheck.txt; cat gpurun_out/n2_memcheck.txt

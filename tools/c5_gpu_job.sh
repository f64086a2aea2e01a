set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_lanczos.py -x -q 2>&1 | grep -v DEBUG | tail -15
timeout 900 python bench.py --config c5 --steps 5 --warmup 3 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
tail -3 gpurun_out/c5_bench.err
cat gpurun_out/c5_bench.json

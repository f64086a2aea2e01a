set -x
mkdir -p gpurun_out
timeout 1500 python bench.py --config c4 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c4_bench_e2e.json 2> gpurun_out/c4_bench.err
tail -5 gpurun_out/c4_bench.err | cut -c1-400
python -c "
import json; d=json.loads(open('gpurun_out/c4_bench_e2e.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e'])"

set -x
mkdir -p gpurun_out
timeout 1500 python bench.py --config c4 --steps 5 --warmup 3 > gpurun_out/c4_bench.json 2> gpurun_out/c4_bench.err
tail -3 gpurun_out/c4_bench.err | cut -c1-300
python -c "
import json; d=json.loads(open('gpurun_out/c4_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['ms_per_step'], d.get('gpu_library_baseline'), d['self_check'], d['first_product_s'])"
CURV_MODEL=vit_b_16 CURV_DTYPE=bf16 CURV_B=32 CURV_K=4 timeout 600 python tools/gpu_step_profile.py 2>&1 | grep -v "DEBUG\|arn" | head -45 > gpurun_out/c4_step_profile.txt

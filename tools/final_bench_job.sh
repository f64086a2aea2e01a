# Final single-GPU measurement pass of a round: every bench configuration with its baselines + step profiles.
set -x
mkdir -p gpurun_out/final
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/final/bench_c2.json 2> gpurun_out/final/bench_c2.err
timeout 600 python bench.py --config c2-bf16 --steps 20 --warmup 3 > gpurun_out/final/bench_c2bf16.json 2> gpurun_out/final/bench_c2bf16.err
timeout 900 python bench.py --config c2-hessian --steps 10 --warmup 3 > gpurun_out/final/bench_c2_hessian.json 2> gpurun_out/final/bench_c2_hessian.err
timeout 600 python bench.py --config c1 > gpurun_out/final/bench_c1.json 2> gpurun_out/final/bench_c1.err
timeout 900 python bench.py --config c3 > gpurun_out/final/bench_c3.json 2> gpurun_out/final/bench_c3.err
timeout 900 python bench.py --config c5 > gpurun_out/final/bench_c5.json 2> gpurun_out/final/bench_c5.err
timeout 900 python bench.py --config c4 > gpurun_out/final/bench_c4.json 2> gpurun_out/final/bench_c4.err
timeout 300 python tools/gpu_step_profile.py 2>&1 | grep -v "DEBUG\|Warn\|warn" > gpurun_out/final/step_profile_fp32.txt
CURV_DTYPE=bf16 timeout 300 python tools/gpu_step_profile.py 2>&1 | grep -v "DEBUG\|Warn\|warn" > gpurun_out/final/step_profile_bf16.txt
CURV_OP=hessian timeout 300 python tools/gpu_step_profile.py 2>&1 | grep -v "DEBUG\|Warn\|warn" > gpurun_out/final/step_profile_hessian.txt
CURV_MODEL=vit_b_16 CURV_DTYPE=bf16 CURV_B=32 CURV_K=4 timeout 300 python tools/gpu_step_profile.py 2>&1 | grep -v "DEBUG\|Warn\|warn" > gpurun_out/final/step_profile_c4.txt
CURV_MODEL=resnet50 CURV_DTYPE=bf16 CURV_B=64 CURV_K=1 timeout 300 python tools/gpu_step_profile.py 2>&1 | grep -v "DEBUG\|Warn\|warn" > gpurun_out/final/step_profile_c5.txt
CURV_B=32 CURV_K=4 timeout 300 python tools/gpu_step_profile.py 2>&1 | grep -v "DEBUG\|Warn\|warn" > gpurun_out/final/step_profile_b32_k4.txt
for f in gpurun_out/final/*.json; do echo "== $f"; cut -c1-330 $f; done
for f in gpurun_out/final/*.err; do tail -n 2 $f | cut -c1-200; done

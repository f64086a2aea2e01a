"""The reference arm of bench.py (the CPU leg the driver runs beside the GPU arm) prints ONE JSON line with the
contract's keys; it needs no GPU.  The arm runs the unmodified reference from oracle/_ref (built here by
oracle/build_ref.py) at full size; this test shrinks the mini-batch through the arm's documented test knob (~20 s)
and launches it the way torchrun would (OMP_NUM_THREADS=1) to check that it still uses every host core."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    sys.path.insert(0, ROOT)
    from oracle.build_ref import build

    assert build() is not None, "oracle/_ref could not be built (no /root/reference and no prebuilt copy)"
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", CURV_BENCH_REF_SAMPLE="4", OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["higher_is_better"] is True and line["vs_baseline"] is None
    assert line["unit"] == "param*vec/s" and line["value"] > 0 and line["ms_per_step"] > 0
    assert "ResNet-18" in line["config"]["workload"] and "model" not in line["config"]
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["sample"] and cb["value"] == line["value"]
    assert cb["cores"] == len(os.sched_getaffinity(0))  # not the single thread torchrun's OMP_NUM_THREADS=1 would give
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}


def test_reference_arm_non_zero_ranks_exit_silently():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="", RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == "", (out.stdout[-500:], out.stderr[-500:])

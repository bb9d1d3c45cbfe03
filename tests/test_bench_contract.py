"""bench.py's reference arm runs on a CPU-only box and prints ONE JSON line with the contract's keys (the CUDA arm needs a GPU and
is exercised on the B200 box)."""
import json
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--no-extras"],
                         cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().split("\n") if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "resnet3d18_train_volumes_per_sec" and d["unit"] == "volumes/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["n_gpus"] == 1 and d["value"] > 0
    assert d["config"]["workload"] == "resnet3d18_bf16_train_batch16_1x128^3_3class"
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "batch 2" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_flop_models_are_consistent():
    sys.path.insert(0, ROOT)
    import bench
    from multimodal_ad_b200.models import resnet

    m = resnet.resnet18(sample_input_D=128, sample_input_H=128, sample_input_W=128, num_seg_classes=1)
    f = bench.resnet_conv_flops_model(m, 16, (128, 128, 128))
    assert abs(f - 14.61e12) < 0.02e12                                  # the figure DESIGN.md quotes
    assert abs(bench.unet3d_forward_flops() - 1.92e12) < 0.03e12

"""bench.py contract checks that need no GPU: the reference arm prints exactly ONE JSON line on stdout with the keys
the driver reads, and the GPU arm's keys are produced from the same code paths (checked structurally)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--cpu-sample-steps", "1", "--envs", "8", "--minibatch", "8", "--epochs", "1"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, "stdout must carry the JSON line only"
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "env-steps/s" and d["higher_is_better"] is True
    assert d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0 and d["value"] > 0 and d["ms_per_step"] > 0
    assert d["vs_baseline"] is None and d["scaling"] == "weak" and d["data"] == "synthetic" and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_flop_constants_match_the_survey_table():
    """per-sample MACs and the per-env-step FLOP figures bench.py divides by"""
    sys.path.insert(0, ROOT)
    import bench
    macs = bench.MACS
    assert (macs["conv0"], macs["conv1"], macs["conv2"], macs["fc"]) == (475 * 32 * 256, 108 * 64 * 512, 108 * 64 * 576, 6912 * 512)
    fwd = 2 * (sum(macs.values()) + 512 * 5)
    assert abs(fwd - 29.906e6) < 2e3
    f, t, cf, ct = bench.FLOPS["gray"]
    assert abs(f - fwd) < 2e3 and abs(t - 81.936e6) < 1e3
    assert abs((f * (1 + 1 / 128) + 4 * t) - 357.9e6) < 1e5 and abs((cf * (1 + 1 / 128) + 4 * ct) - 265.7e6) < 1e5
    n = bench.MACS_NATURE84
    assert (n["conv0"], n["conv1"], n["conv2"], n["fc"]) == (400 * 32 * 256, 81 * 64 * 512, 49 * 64 * 576, 3136 * 512)
    f84, t84, _, _ = bench.FLOPS["rgb"]
    assert abs(2 * (sum(n.values()) + 512 * 5) - f84) < 2e4 and abs((f84 * (1 + 1 / 128) + 4 * t84) - 216.9e6) < 1e5

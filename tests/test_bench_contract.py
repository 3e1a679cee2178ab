"""bench.py's output contract: the reference arm (`--impl reference`, the oracle port on the host cores) runs without a
GPU and prints the agreed JSON line; the roofline numerator of the table kernel is what DESIGN.md says it is.  (The GPU arm's
line is checked where a GPU exists: tests/test_bench_gpu.py.)"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _check_common(d):
    assert BASE_KEYS <= set(d), BASE_KEYS - set(d)
    assert d["metric"] == "decoded syndromes/sec" and d["unit"] == "syndromes/s" and d["higher_is_better"] is True
    assert d["scaling"] == "weak" and d["data"] == "synthetic" and "workload" in d["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["ms_per_step"] > 0


def test_reference_arm_runs_on_cpu_and_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    _check_common(d)
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["config"]["workload"] == "v2_4_rotated_d5_depol_B65536"        # BASELINE.json configs[1]


def test_wavefront_count_of_the_table_kernel():
    """bench.py's algorithmic shared-memory wavefront count (the roofline numerator of the check-owner table kernel) on the
    headline code: rotated d = 5 has 80 edges, 60 of them on variables with a second edge, 24 checks, 50 variables."""
    sys.path.insert(0, ROOT)
    import bench
    from gnn_decode_b200 import codes
    pcm = codes.rotated_surface_pcm(5)
    assert bench.lean_wavefronts_per_syndrome(pcm, 15) == (14 * (6 * 80 + 5 * 60) + (80 + 4 * 24) + 5 * 80 + 2 * 50) / 32.0

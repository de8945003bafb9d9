"""bench.py prints exactly one JSON line with the keys the driver reads (GPU; a reduced source count keeps it short)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_line_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "1", "--nsrc", "600",
                          "--no-eor", "--cpu-sample", "16"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "roofline", "cpu_baseline", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["steps"] == 2 and d["dtype"] == "f64" and d["higher_is_better"] is True
    assert d["value"] > 1e9 and d["e2e"]["value"] > 1e8 and d["e2e"]["value"] <= d["value"] * 1.05
    assert d["e2e"]["h2d_bytes_per_step"] == 8 * 256 ** 3 == d["e2e"]["d2h_bytes_per_step"]
    r = d["roofline"]
    assert r["bound"] == "hbm" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12 and r["unit"] == "GB/s"
    assert d["gpu_launches"] >= 2 * d["steps"]
    par = d["parity"]
    assert par["max_rel"] <= par["tolerance"] == 1e-9 and par["cells"] > 1000 and "failed_checks" not in d
    assert d["e2e_pinned"]["value"] >= 0.9 * d["e2e"]["value"] and "pageable" in d["e2e"]["host_buffers"]
    rg = d["reference_gpu"]
    assert rg["batch64"]["kernel_updates_per_s"] > 1e8 and rg["batch128"]["e2e_ms"] > 0
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["single_thread"]["value"] > 0
    assert "workload" in d["config"]


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "16"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["unit"] == "updates/s"
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["config"]["cells_swept_per_source"] >= d["config"]["cells_credited_per_source"]


def test_reference_arm_ignores_torchrun_thread_cap_and_does_not_load_the_cuda_library():
    """torchrun exports OMP_NUM_THREADS=1; the arm must still use the host's cores, finish in bounded time, and leave
    libasora_b200.so unloaded."""
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0', '--cpu-sample', '64'];"
            "runpy.run_path(%r, run_name='__main__');"
            "maps = open('/proc/self/maps').read(); assert 'libasora_b200' not in maps, 'CUDA library loaded'" %
            os.path.join(ROOT, "bench.py"))
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))

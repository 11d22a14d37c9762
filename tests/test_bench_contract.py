"""The bench.py output contract, checked on the CPU arm (the only arm that runs without a GPU): exactly
one JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 1 and d["steps"] == 1 and d["warmup"] == 0
    for key in ("metric", "value", "unit", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d, key
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"] and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_does_not_map_the_cuda_library():
    """The CPU arm must be the oracle port alone: importing what it imports (the clip generator, the oracle) must not
    dlopen libelvis_b200.so (round-1 verdict: the driver's native-library record for the reference arm was tainted)."""
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import bench\n"
            "from elvis_b200.synth import synth_yuv420\n"
            "from oracle.cpu_baseline import CpuElvisV1, CpuV2\n"
            "from oracle import verify\n"
            "clip = synth_yuv420(2, 32, 64, seed=1, device='cpu')\n"
            "assert 'libelvis_b200' not in open('/proc/self/maps').read(), 'the CUDA library was mapped'\n"
            "assert not any(m.startswith('elvis_b200.') and m.split('.')[1] in ('ops', '_lib', 'pipeline') for m in sys.modules)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]


def test_bench_helpers():
    sys.path.insert(0, ROOT)
    import numpy as np
    import bench
    b = bench.v1_bytes_per_frame(3840, 2160)
    assert b["total"] == 39398400 and bench.v2_bytes_per_frame(3840, 2160, False) == 3 * 3840 * 2160
    s = np.array([[0.0, 0.125, 0.5, 0.874, 0.876, 1.0]])
    assert bench.v2_map_from_scores("downsample", s).tolist() == [[0, 0, 2, 3, 3, 3]]                     # round-half-even, clamped to 2 bits
    assert bench.v2_map_from_scores("blur", s).tolist() == [[0, 1, 5, 9, 9, 10]]
    assert bench.v2_map_from_scores("dampen", s).dtype == np.float32
    tr = bench.ncu_traffic("score_umma_kernel")
    assert tr and tr["frames"] == 120 and 0.99e9 < tr["bytes"] < 1.1e9 and "profiles/" in tr["source"]
    assert bench.ncu_traffic("no_such_kernel") is None
    for wl in bench.SPECS:
        class A:
            clips_per_step, distinct = 100, 4
        cfg = bench.config_dict(wl, 2, 120, A)
        assert "workload" in cfg and "model" not in cfg and cfg["dct_size"] == 8

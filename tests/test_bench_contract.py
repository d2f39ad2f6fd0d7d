"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*extra):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        *extra], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.strip().splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_json_contract():
    d = _run("--ref-sweeps", "1")
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "chain-steps/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["value"] > 0 and "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_lean_c_arm_is_much_faster_than_faithful():
    slow = _run("--ref-sweeps", "1")["value"]
    fast = _run("--ref-sweeps", "200", "--ref-mode", "lean")["value"]
    assert fast > 20 * slow


def test_non_zero_ranks_of_reference_arm_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_roofline_traffic_comes_from_the_committed_ncu_capture():
    """bench.py's roofline.traffic is read from profiles/r*_ncu_sweep_dram_bench.csv (no hard-coded constant) and only
    when that capture is of the same kernel, workload and (within 25 %) launch time."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    b = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(b)
    traffic, note = b._ncu_dram_traffic(131072, 100, 1180.0)
    assert traffic is not None and 1.0e10 < traffic < 3.0e10 and "mh_sweep_kernel" in note
    assert b._ncu_dram_traffic(131072, 100, 2500.0)[0] is None          # another kernel generation: launch time far off
    assert b._ncu_dram_traffic(4096, 100, 1180.0)[0] is None            # another workload
    assert b._ncu_dram_traffic(131072, 50, 1180.0)[0] is None

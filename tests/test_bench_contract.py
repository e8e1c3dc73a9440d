"""The bench.py output contract, checked on the bench lines committed under profiles/ (they are what the last GPU runs of
the round printed): every key the driver and the judge read is present and well-formed.  bench.py itself needs a GPU; its
CPU-only `--impl reference` arm is exercised here on a tiny sample."""
import glob
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def lines(pattern):
    out = []
    for f in sorted(glob.glob(os.path.join(ROOT, "profiles", pattern))):
        with open(f) as fh:
            out.append((os.path.basename(f), json.loads(fh.read().strip().splitlines()[-1])))
    return out


@pytest.mark.parametrize("name,d", lines("r01_final_b_c*.json") + lines("r01_final_b_app6.json") + lines("r01_b_c2_n*.json") +
                         lines("r02_final_b_n1.json") + lines("r02_bench_c2_n8.json"))
def test_committed_bench_lines_follow_the_contract(name, d):
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "e2e", "gpu_launches", "clocks", "roofline"):
        assert k in d, "%s: missing %s" % (name, k)
    assert d["unit"] == "frames/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["data"] == "synthetic"
    assert d["warmup"] >= 3 and d["value"] > 0 and d["gpu_launches"] > 0
    assert "workload" in d["config"] and "model" not in d["config"] and "l2_policy" in d["config"]
    e = d["e2e"]
    assert e["value"] > 0 and e["unit"] == "frames/s" and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert e["value"] < d["value"]                                     # host buffers can only be slower than resident inputs
    c = d["clocks"]
    assert c["sm_mhz"] and c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    assert abs(r["achieved"] - r["algorithmic_bytes_per_launch"] / (r["avg_launch_us"] * 1e-6) / 1e9) < 1e-3 * r["achieved"]
    if d["n_gpus"] == 1:
        b = d["cpu_baseline"]
        assert b["value"] > 0 and b["cores"] >= 1 and b["kind"] in ("reference", "port") and b["sample"]
    if "app6" not in name:
        assert d["vs_baseline"] is None                                # BASELINE.md publishes nothing for these workloads
    else:
        assert d["vs_baseline"] == pytest.approx(d["value"] / (1000.0 / 43.6))


def test_round2_line_carries_the_heavy_workloads():
    """the default N=1 line reports C3 and the live app's case measured in the same run, the copy-only PCIe
    ceiling next to e2e, a timed region of ~0.5 s, and the kernel the roofline names is the TMA frame kernel."""
    name, d = lines("r02_final_b_n1.json")[0]
    assert set(d["workloads"]) == {"c3", "app6"}
    for w, m in d["workloads"].items():
        for k in ("metric", "value", "unit", "ms_per_step", "config", "e2e", "gpu_launches", "clocks", "roofline", "kernels"):
            assert k in m, "%s: missing %s" % (w, k)
        assert m["value"] > 0 and m["e2e"]["value"] > 0 and 0 < m["roofline"]["frac"] < 1 and m["gpu_launches"] > 0
    assert d["workloads"]["app6"]["vs_baseline"] == pytest.approx(d["workloads"]["app6"]["value"] / (1000.0 / 43.6))
    assert d["steps"] * d["ms_per_step"] >= 400.0                      # timed region >= 0.4 s of device time
    assert d["e2e"]["copy_only_ceiling"]["value"] >= d["e2e"]["value"] * 0.95
    assert d["roofline"]["kernel"] == "feather_stream" and d["roofline"]["traffic"]


@pytest.mark.parametrize("name", ["r01_final_b_ref_c2.json", "r02_final_b_ref.json"])
def test_reference_arm_line(name):
    name, d = lines(name)[0]
    assert d["impl"] == "reference" and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["cores"] >= 1


def test_reference_arm_runs_without_a_gpu():
    """`bench.py --impl reference` is CPU-only: one bounded step of the default workload, same keys."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["metric"] and d["config"]["workload"].startswith("C2")
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["d2h_bytes_per_step"] == 0

"""One line per workload of a bench.py JSON line: python scripts/show_bench.py file.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])


def show(name, m):
    r = m["roofline"]
    per = " = %d frame sets" % r["frame_sets_per_launch"] if "frame_sets_per_launch" in r else ""
    fmt = ("%-5s value %.0f e2e %.0f | dom %s frac %.3f (launch %.1f us" + per + ", isolated %.1f us frac %.3f) | timed %.3f (%.1f us/frame) "
           "single-stream %.3f (%.1f us) | launches %d")
    print(fmt % (name, m["value"], m["e2e"]["value"], r["kernel"], r["frac"], r["avg_launch_us"], r["isolated_launch_us"], r["frac_isolated"],
                 r["timed_region"]["frac"], r["timed_region"]["us_per_frame"], r["single_stream"]["frac"], r["single_stream"]["us_per_frame"],
                 m["gpu_launches"]))
    for k in m.get("kernels", [])[:8]:
        print("        %-18s x%.0f  %.1f MB  in-stream %.1f us  frac %.3f (isolated %.3f)" % (
            k["name"], k["launches_per_frame"], k["algorithmic_mb_per_frame"], k["in_stream_ms_per_frame"] * 1e3, k["frac"], k["frac_isolated"]))


show(d["config"]["workload"][:4], d)
for k, v in d.get("workloads", {}).items():
    show(k, v)

import json, sys
d = json.load(open(sys.argv[1]))
print("value %.3e e2e %.3e one %.3e frac %.3f" % (d["value"], d["e2e"]["value"], d["one_batch_at_a_time"]["value"], d["roofline"]["whole_step"]["frac"]))
print("one_shot", d["one_shot_4096"])
print("pass  ", {k: round(v["ms"], 4) for k, v in d["roofline"]["per_kernel"].items()}, d["roofline"]["one_pass_alone"])
print("job   ", {k: round(v["ms"], 4) for k, v in d["roofline"]["per_kernel_one_job_alone"].items()})
print("e2e   ", {k: v for k, v in d["e2e"].items()})
print("cpu   ", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
ex = d["extras"]
print("extras", ex.get("error"), ex.get("prove", {}).get("value"))
for k, v in ex.get("msm", {}).items():
    print("  msm", k, round(v["mpoints_per_s"], 1), "Mpts/s c=%d" % v["window_bits"], v["phase_ms"], "sort GB/s %.0f" % v["sort_phase"]["gb_per_s"])
for k, v in ex.get("msm_scalar_distributions", {}).get("results", {}).items():
    print("  dist", k, round(v["mpoints_per_s"], 1), v["window_bits"])
print("  sharded", ex.get("msm_sharded"))
print("clocks", d["clocks"], "launches", d["gpu_launches"], "wall", d["wall_s_timed_region"])

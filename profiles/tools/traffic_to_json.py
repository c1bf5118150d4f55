"""gpurun_out/traffic_<name>.csv (ncu --csv of the k_value_mma launches of two steps) -> profiles/value_kernel_traffic.json"""
import csv, json, os, sys, collections
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = {}
for fn in sorted(os.listdir(os.path.join(root, "gpurun_out"))):
    if not (fn.startswith("traffic_") and fn.endswith(".csv")):
        continue
    name = fn[len("traffic_"):-4]
    rows = [r for r in csv.reader(open(os.path.join(root, "gpurun_out", fn))) if len(r) > 10]
    if not rows:
        continue
    h = rows[0]
    iid, ik, im, iv, iu = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    launches = collections.OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(r[iid], {"kernel": r[ik]})
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1, "ms": 1e3, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}.get(u, 1)
        d[r[im]] = v * scale
    L = list(launches.values())
    half = L[len(L) // 2:]  # the second of the two steps the script runs
    big = max(half, key=lambda d: d["gpu__time_duration.sum"])
    out[name] = {
        "dram_bytes_per_launch": int(big["dram__bytes_read.sum"] + big["dram__bytes_write.sum"]),
        "dram_bytes_read": int(big["dram__bytes_read.sum"]), "dram_bytes_write": int(big["dram__bytes_write.sum"]),
        "launch": big["kernel"], "launch_us_under_ncu": big["gpu__time_duration.sum"],
        "launches_per_step": len(half),
        "step_dram_bytes_all_value_launches": int(sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in half)),
        "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_value_mma "
                  "python profiles/exp/prof_run.py (profiles/tools/capture_value_traffic.sh); the longest k_value_mma launch of the second step",
    }
json.dump(out, open(os.path.join(root, "profiles", "value_kernel_traffic.json"), "w"), indent=1)
print(json.dumps(out, indent=1))

"""profiles/traffic.json from an ncu --set full report: DRAM bytes (read + write) per launch and per frame
for each pipeline kernel.  usage: make_traffic.py report.ncu-rep FRAMES_PER_LAUNCH"""
import csv, json, os, subprocess, sys
rep, frames = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, units = rows[0], rows[1]
def num(r, key):
    v = float(r[h.index(key)].replace(",", ""))
    u = units[h.index(key)]
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ms": 1, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1)
res = {}
for r in rows[2:]:
    name = r[h.index("Kernel Name")].split("(")[0].split("::")[-1].split("<")[0].replace("void ", "").strip()
    b = num(r, "dram__bytes_read.sum") + num(r, "dram__bytes_write.sum")
    e = res.setdefault(name, {"bytes_per_launch": 0.0, "frames_per_launch": frames, "duration_ms_under_ncu": 0.0,
                              "launches_summed": 0, "source": os.path.basename(rep)})
    e["bytes_per_launch"] += b; e["duration_ms_under_ncu"] += num(r, "gpu__time_duration.sum"); e["launches_summed"] += 1
    e["bytes_per_frame"] = e["bytes_per_launch"] / frames
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(res, indent=1))

"""profiles/r02_traffic.json from ncu --set full captures of ONE full-size launch (4096 problems) per kernel:
dram__bytes_read.sum + dram__bytes_write.sum divided by the problems of the launch.  bench.py reads the JSON for
`roofline.traffic` (a number from a committed capture, not a literal).
usage: ncu_traffic.py <problems per launch> name=capture.ncu-rep ... > profiles/r02_traffic.json"""
import csv, json, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
n = int(sys.argv[1])
out = {}
for arg in sys.argv[2:]:
    name, rep = arg.split("=")
    rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
    h, u, v = rows[0], rows[1], rows[2]
    tot = 0.0
    for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        i = h.index(k)
        tot += float(v[i]) * UNIT[u[i]]
    out[name] = {"capture": rep.split("/")[-1], "kernel": v[h.index("Kernel Name")].split("(")[0], "problems": n,
                 "gpu_time_ms": float(v[h.index("gpu__time_duration.sum")]) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6}[u[h.index("gpu__time_duration.sum")]],
                 "dram_bytes_per_problem": tot / n}
json.dump(out, sys.stdout, indent=1)

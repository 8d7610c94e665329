#!/usr/bin/env python
"""ncu csv (--metrics dram__bytes_read.sum,dram__bytes_write.sum[,gpu__time_duration.sum] --csv) of ONE launch ->
profiles/traffic_<kernel>_<tag>.json, the file bench.py's roofline.traffic quotes.  The JSON records the hash of the kernel
sources it was captured from: bench.py does not quote a capture of other sources.
usage: traffic_json.py <ncu.csv> <kernel> <workload> <n_steps> <n_burn> <tag>"""
import csv, datetime, hashlib, json, os, sys
path, kernel, workload, n_steps, n_burn, tag = sys.argv[1:7]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
h = hashlib.sha256()
for f in ("tc_mcmc.cu", "tc_device.cuh", "tc_warp.cuh"):
    h.update(open(os.path.join(ROOT, "transcriptioncycleinference_b200", "csrc", f), "rb").read())
rows = [r for r in csv.reader(open(path)) if r]
hdr = next(r for r in rows if r[0] == "ID")
vals = {}
for r in rows:
    if r[0].isdigit():
        d = dict(zip(hdr, r))
        if kernel in d["Kernel Name"]:
            v = float(d["Metric Value"].replace(",", ""))
            u = d["Metric Unit"]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}.get(u, 1)
            vals[d["Metric Name"]] = v * scale
out = dict(kernel=kernel, workload=workload, n_steps=int(n_steps), n_burn=int(n_burn), tag=tag,
           bytes_read=int(vals["dram__bytes_read.sum"]), bytes_write=int(vals["dram__bytes_write.sum"]),
           seconds_under_ncu=vals.get("gpu__time_duration.sum"), source_hash=h.hexdigest()[:16],
           when=datetime.datetime.utcnow().strftime("%Y-%m-%dT%H:%M:%SZ"),
           how="ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none, one launch")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
p = os.path.join(ROOT, "gpurun_out", "traffic_%s_%s.json" % (kernel, tag))
json.dump(out, open(p, "w"), indent=1)
print(p, out)

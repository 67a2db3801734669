#!/bin/bash
# DRAM traffic of ONE launch of the headline kernel over the bench corpus (1 GiB), for
# bench.py's roofline.traffic.  Writes gpurun_out/traffic.json stamped with the hash of the
# kernel source it was measured on; copy it to profiles/traffic.json.
set -u
mkdir -p gpurun_out
O=gpurun_out
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:scanKernelPair -s 4 -c 1 --csv --log-file $O/traffic_ncu.csv \
    python bench.py --steps 1 --warmup 1 --passes-per-step 4 --no-e2e --no-cpu --no-secondary > $O/traffic_bench.out 2>&1
python - <<'PY'
import csv, hashlib, json
rows = [r for r in csv.reader(open("gpurun_out/traffic_ncu.csv", errors="ignore")) if len(r) > 10]
hdr = rows[0]
vals = {}
kern = ""
for r in rows[1:]:
    d = dict(zip(hdr, r))
    kern = d.get("Kernel Name", kern)
    vals[d["Metric Name"]] = float(d["Metric Value"].replace(",", ""))
sha = hashlib.sha256(open("hyperscan_b200/csrc/device/scan_kernels.cu", "rb").read()).hexdigest()[:16]
out = {"kernel": kern, "dram_bytes_per_launch": int(vals["dram__bytes_read.sum"] + vals["dram__bytes_write.sum"]),
       "dram_bytes_read": int(vals["dram__bytes_read.sum"]), "dram_bytes_write": int(vals["dram__bytes_write.sum"]),
       "duration_ns_under_ncu": vals.get("gpu__time_duration.sum"), "corpus_bytes": 1 << 30,
       "kernel_source_sha16": sha,
       "command": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:scanKernelPair -s 4 -c 1 "
                  "python bench.py --steps 1 --warmup 1 --passes-per-step 4 --no-e2e --no-cpu --no-secondary"}
json.dump(out, open("gpurun_out/traffic.json", "w"), indent=1)
print(out)
PY

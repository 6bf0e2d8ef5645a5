"""gpurun_out/r2_gemm_traffic.csv (ncu dram bytes + duration of the step's gemm_tc launches) -> profiles/r02_gemm_dram_traffic.json"""
import csv
import json
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(r for r in rows if "Kernel Name" in r)
data = [dict(zip(hdr, r)) for r in rows if len(r) == len(hdr) and r is not hdr]
per = {}
for d in data:
    per.setdefault(d["ID"], {})[d["Metric Name"]] = (float(d["Metric Value"].replace(",", "")), d["Metric Unit"])


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def to_us(v, u):
    return v * {"ns": 1e-3, "us": 1, "ms": 1e3}.get(u, 1e-3)


rd = [to_bytes(*m["dram__bytes_read.sum"]) for m in per.values()]
wr = [to_bytes(*m["dram__bytes_write.sum"]) for m in per.values()]
us = [to_us(*m["gpu__time_duration.sum"]) for m in per.values()]
n = len(rd)
out = {
    "kernel": "gemm_tc_kernel (150 launches = two training steps, bf16 mode, N=1, 64x256 tokens)",
    "avg_dram_bytes_per_launch": (sum(rd) + sum(wr)) / n, "avg_dram_read_bytes": sum(rd) / n, "avg_dram_write_bytes": sum(wr) / n,
    "launches_measured": n, "avg_time_us_under_ncu": sum(us) / n,
    "source": "profiles/r02_gemm_dram_traffic.csv: ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
              "--clock-control none -k regex:gemm_tc --launch-skip 300 -c 150 python bench.py --steps 2 --warmup 3 --no-parity --no-fp32 "
              "--no-cpu-baseline (caches flushed between launches by ncu: cold-cache reads; writes mostly stay in the 126 MB L2)",
}
json.dump(out, open(sys.argv[2], "w"), indent=1)
print(out["avg_dram_bytes_per_launch"], n)

"""Reads `ncu --set full` captures (gpurun_out/*.ncu-rep) and writes profiles/ncu_traffic.json: per captured launch the measured DRAM
traffic (dram__bytes_read.sum + dram__bytes_write.sum), duration and tensor-pipe activity.  bench.py reads `dram_bytes` from this file
for `roofline.traffic` (never a typed-in constant).     python tools/ncu_traffic.py key=path.ncu-rep [key=path ...]"""
import csv
import io
import json
import os
import subprocess
import sys

UNITS = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def read(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    g = {h: (u, v) for h, u, v in zip(hdr, units, vals)}

    def num(name, scale_table=True):
        u, v = g[name]
        x = float(v.replace(",", ""))
        return x * UNITS.get(u, 1) if scale_table else x
    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    return dict(kernel=g["Kernel Name"][1], dram_bytes_read=int(rd), dram_bytes_write=int(wr), dram_bytes=int(rd + wr),
                duration_us=round(num("gpu__time_duration.sum"), 2),
                tensor_pipe_active_pct=round(num("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False), 2),
                sm_ghz=round(num("sm__cycles_elapsed.avg.per_second", False), 3), src=os.path.basename(path))


if __name__ == "__main__":
    res = {}
    p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        res = json.load(open(p))
    for arg in sys.argv[1:]:
        k, path = arg.split("=", 1)
        res[k] = read(path)
        print(k, res[k])
    json.dump(res, open(p, "w"), indent=1)

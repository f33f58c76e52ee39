"""`python tools/ncu_summary.py name=path.ncu-rep ...` -> text summary (key metrics + warp-stall mix + top stall sites) for profiles/."""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]


def page(path, which):
    return subprocess.run(["ncu", "-i", path, "--page", which, "--csv"], capture_output=True, text=True, check=True).stdout


for arg in sys.argv[1:]:
    name, path = arg.split("=", 1)
    rows = list(csv.reader(io.StringIO(page(path, "raw"))))
    g = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
    print(f"== {name} : {g['Kernel Name'][1][:110]}")
    for k in KEYS:
        if k in g:
            print(f"   {k} = {g[k][1]} {g[k][0]}")
    st = sorted(((float(v[1]), k.split('issue_stalled_')[1].split('_per_')[0]) for k, v in g.items()
                 if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and v[1]), reverse=True)
    print("   warp stalls per issue: " + ", ".join(f"{n}={x:.2f}" for x, n in st[:7]))
    try:
        src = list(csv.reader(io.StringIO(page(path, "source"))))
        hdr = src[1]
        ia, isrc, ins = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples")
        rows2 = []
        for r in src[2:]:
            try:
                rows2.append((int(r[ins]), r[isrc].strip()[:70]))
            except (ValueError, IndexError):
                pass
        tot = sum(n for n, _ in rows2) or 1
        rows2.sort(reverse=True)
        print(f"   top sampled SASS sites ({tot} samples): " + "; ".join(f"{100 * n / tot:.0f}% {t}" for n, t in rows2[:6]))
    except Exception as e:  # noqa: BLE001
        print("   (no source page:", e, ")")

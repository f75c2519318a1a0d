"""Digest of an .ncu-rep: headline metrics (raw page) and the top stall lines (source page).
usage: python tools/ncu_digest.py file.ncu-rep [kernel-regex]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
def page(p):
    return subprocess.run(["ncu", "-i", rep, "--page", p, "--csv"], capture_output=True, text=True).stdout
raw = list(csv.reader(io.StringIO(page("raw"))))
h = raw[0]
want = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum"]
for row in raw[2:]:
    for w in want:
        for i, x in enumerate(h):
            if x == w:
                print(f"{w:75s} {row[i]} {raw[1][i]}")
    print("---")
src = list(csv.reader(io.StringIO(page("source"))))
hdr = [i for i, r in enumerate(src) if r and r[0] == "Address"]
if hdr:
    hh = src[hdr[0]]
    body = src[hdr[0] + 1:(hdr[1] - 1 if len(hdr) > 1 else len(src))]
    si, sc = hh.index("# Samples"), hh.index("Source")
    stalls = [i for i, x in enumerate(hh) if x.startswith("stall_") and "Not Issued" not in x]
    body = [r for r in body if len(r) > si]
    tot = sum(int(r[si] or 0) for r in body)
    agg = {}
    for r in body:
        for i in stalls:
            agg[hh[i]] = agg.get(hh[i], 0) + int(r[i] or 0)
    print("samples", tot, "sass lines", len(body))
    print("  ".join(f"{k[6:]}={v}" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
    for r in sorted(body, key=lambda r: -int(r[si] or 0))[:14]:
        d = {hh[i][6:]: int(r[i] or 0) for i in stalls if int(r[i] or 0) > 0}
        print(f"{r[si]:>6s} {r[sc][:80]:80s} {dict(sorted(d.items(), key=lambda kv: -kv[1])[:3])}")

"""Per-CUDA-source-line stall samples of an .ncu-rep (needs -lineinfo + --import-source on).
usage: python tools/ncu_lines.py file.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 16
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
cur, hdr, out = None, None, []
for r in csv.reader(io.StringIO(txt)):
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r and r[0] == "Line No":
        hdr = r
        si = hdr.index("# Samples")
    elif hdr and len(r) > si and r[0].strip().isdigit():
        try:
            out.append((int(r[si]), cur, r[0], " ".join(r[1].split())[:110]))
        except ValueError:
            pass
tot = sum(o[0] for o in out)
print("samples", tot)
for n, f, l, s in sorted(out, reverse=True)[:top]:
    print(f"{n:6d} {100.0 * n / max(tot, 1):5.1f}%  {f}:{l}  {s}")

"""Aggregate an ncu report's warp-stall samples by CUDA source line.
usage: python scripts/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = None; hdr = None; items = []; tot = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[2] != "-": continue          # only the per-source-line summary rows
    try: n = int(r[hdr.index("# Samples")])
    except ValueError: continue
    if n <= 0: continue
    stalls = sorted(((int(r[i]) if r[i].isdigit() else 0, c) for i, c in enumerate(hdr) if c.startswith("stall_") and "Not Issued" not in c), reverse=True)[:3]
    inst = r[hdr.index("Instructions Executed")]
    items.append((n, cur_file, r[0], r[1].strip()[:90], stalls, inst)); tot += n
items.sort(reverse=True)
print("total samples", tot)
for n, f, ln, src, st, inst in items[:top]:
    print(f"{100*n/tot:5.1f}% {f}:{ln:>4} inst={inst:>10} {src}   {[(c[6:], k) for k, c in st]}")

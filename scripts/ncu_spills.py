"""Aggregate executed local-memory (register spill) instructions of an ncu report by CUDA source line.
usage: python scripts/ncu_spills.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = None; hdr = None; cur_line = None; cur_src = ""
agg = collections.Counter(); src = {}; total_exec = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    if r[2] == "-":
        cur_line = (cur_file, r[0]); src[cur_line] = r[1].strip()[:80]; continue
    # SASS row: r[1] is the instruction text
    try: n = int(r[hdr.index("Instructions Executed")])
    except ValueError: continue
    total_exec += n
    ins = r[3]
    if "LDL" in ins or "STL" in ins: agg[cur_line] += n
tot = sum(agg.values())
print(f"local-memory instructions executed: {tot} of {total_exec} ({100*tot/max(total_exec,1):.1f} %)")
for k, n in agg.most_common(top):
    print(f"{100*n/tot:5.1f}% {k[0]}:{k[1]:>4} {n:>12} {src.get(k, '')}")

"""Aggregate an ncu report's warp-stall samples by source REGION (function-level ranges given on the
command line as file:lo-hi=name ...).  usage: python scripts/ncu_regions.py report.ncu-rep regions.txt"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
regions = []
for ln in open(sys.argv[2]):
    ln = ln.strip()
    if not ln or ln.startswith("#"): continue
    spec, name = ln.split("=")
    f, rng = spec.split(":"); lo, hi = rng.split("-")
    regions.append((f, int(lo), int(hi), name))
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file = None; hdr = None
agg = collections.defaultdict(lambda: collections.Counter()); tot = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; hdr = None; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[2] != "-": continue
    try: n = int(r[hdr.index("# Samples")]); line = int(r[0])
    except ValueError: continue
    if n <= 0: continue
    name = f"{cur_file}:other"
    for f, lo, hi, nm in regions:
        if f == cur_file and lo <= line <= hi: name = nm; break
    agg[name]["samples"] += n; tot += n
    for i, c in enumerate(hdr):
        if c.startswith("stall_") and "Not Issued" not in c and r[i].isdigit(): agg[name][c[6:]] += int(r[i])
print("total samples", tot)
for name, c in sorted(agg.items(), key=lambda kv: -kv[1]["samples"]):
    st = [(k, v) for k, v in c.most_common(5) if k != "samples"][:4]
    print(f"{100*c['samples']/tot:5.1f}%  {name:28s} {st}")

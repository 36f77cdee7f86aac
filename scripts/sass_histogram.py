"""Opcode histogram of every kernel in libpareben.so (cuobjdump -sass), written to profiles/<tag>_sass_histogram.md.
usage: python scripts/sass_histogram.py r02      (runs here: no GPU needed)"""
import collections, os, re, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "pareben_b200", "libpareben.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern.replace("(anonymous namespace)::", "")).replace("pareben::", "").replace("void ", "")
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_]+)*)", line)
    if m and kern:
        op = m.group(1)
        hist[kern][op.split(".")[0]] += 1
        if op.startswith(("DMMA", "UBLKCP", "SYNCS", "LDGSTS", "UTMALDG", "I2F.F64")):
            hist[kern]["=" + ".".join(op.split(".")[:3])] += 1
keys = ["DMMA", "DFMA", "DMUL", "DADD", "UBLKCP", "SYNCS", "LDGSTS", "UTMALDG", "LDG", "LDS", "STS", "STG", "LDL", "STL", "I2F", "SHFL", "BAR", "MUFU"]
with open(os.path.join(ROOT, "profiles", f"{tag}_sass_histogram.md"), "w") as f:
    f.write(f"# {tag} -- SASS opcode histogram of `pareben_b200/libpareben.so` (sm_100a)\n\n"
            "`cuobjdump -sass`, static instruction counts per kernel (`python scripts/sass_histogram.py`).  `DMMA` = `mma.sync.m8n8k4.f64` "
            "(the FP64 tensor path; tcgen05 has no f64 kind), `UBLKCP` = `cp.async.bulk` executed by the TMA unit, `SYNCS` = mbarrier "
            "operations, `LDGSTS` = `cp.async`, `LDL`/`STL` = local memory (spills).\n\n| kernel | total | " + " | ".join(keys) + " |\n|---|---|" + "---|" * len(keys) + "\n")
    for k, h in hist.items():
        tot = sum(v for o, v in h.items() if not o.startswith("="))
        if tot < 50:
            continue
        f.write(f"| `{k}` | {tot} | " + " | ".join(str(h.get(o, 0)) for o in keys) + " |\n")
    f.write("\nVariants seen:\n\n")
    for k, h in hist.items():
        v = sorted((o[1:], n) for o, n in h.items() if o.startswith("="))
        if v:
            f.write(f"* `{k}`: " + ", ".join(f"`{o}` x{n}" for o, n in v) + "\n")
print(open(os.path.join(ROOT, "profiles", f"{tag}_sass_histogram.md")).read()[:3000])

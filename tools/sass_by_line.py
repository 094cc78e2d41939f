"""Joins an `ncu --page source --csv` SASS listing with `nvdisasm -g -c` line info of the same kernel:
instructions executed and stall samples per CUDA source line.

  cuobjdump -xelf all lib.so; nvdisasm -g -c features.sm_100a.cubin > feat.dis
  ncu -i rep.ncu-rep --page source --csv > src.csv
  python tools/sass_by_line.py src.csv feat.dis <kernel-substring> <source.cu> [per-unit]
"""
import csv, re, sys, collections

src_csv, dis, kname, cu = sys.argv[1:5]
unit = float(sys.argv[5]) if len(sys.argv) > 5 else 1.0
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]; data = rows[2:]
ia = hdr.index("Instructions Executed"); isamp = hdr.index("# Samples"); isrc = hdr.index("Source")
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kname in l)
cur = None; seq = []
for l in lines[start + 1:]:
    if l.startswith("//---") and ".text." in l: break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        # inlined frames: keep the outermost location inside the .cu file when present
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m: seq.append((cur, m.group(2)))
assert len(seq) == len(data), (len(seq), len(data))
agg = collections.defaultdict(lambda: [0, 0])
for (loc, txt), r in zip(seq, data):
    key = loc if loc else ("?", 0)
    agg[key][0] += int(r[ia]); agg[key][1] += int(r[isamp])
text = open(cu).read().split("\n")
tot = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
print("total inst/unit %.1f  samples %d" % (tot / unit, ts))
for (f, ln), (n, s) in sorted(agg.items(), key=lambda kv: (kv[0][0], kv[0][1])):
    if n / tot < 0.004 and s / max(ts, 1) < 0.004: continue
    t = text[ln - 1].strip()[:90] if f.endswith(cu.split("/")[-1]) and 0 < ln <= len(text) else f.split("/")[-1]
    print("%5d %8.2f %6.1f%%  %s" % (ln, n / unit, 100.0 * s / max(ts, 1), t))

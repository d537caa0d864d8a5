"""Attribute an ncu capture's SASS-level samples / executed instructions to CUDA source lines.
    python tools/ncu_lines.py <report.ncu-rep> <lib.so> <kernel mangled-name substring> [top]
Joins `ncu --page source --csv` (one row per SASS instruction, in order) with `nvdisasm -g` line annotations."""
import csv, os, subprocess, sys, tempfile, collections, io

rep, lib, pat = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; data = rows[2:]
isamp, iex, ith, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed"), hdr.index("Source")
with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
    cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], capture_output=True, text=True).stdout.splitlines()
lines = []; cur = ("?", 0); inside = False; seen = False
for l in dis:
    if l.startswith("//---") and ".text." in l:
        inside = pat in l and not seen          # several instantiations may match the pattern: the first one (give a longer pattern to pick another)
        seen = seen or inside
        continue
    if not inside:
        continue
    s = l.strip()
    if s.startswith('//## File'):
        f = s.split('"')[1]; n = int(s.rsplit("line", 1)[1].split()[0]); cur = (os.path.basename(f), n)
    elif s.startswith("/*") and "*/" in s and not s.startswith("/* "):
        lines.append(cur)
assert len(lines) == len(data), (len(lines), len(data))
agg = collections.defaultdict(lambda: [0, 0, 0.0])
for (f, n), r in zip(lines, data):
    a = agg[(f, n)]; a[0] += int(r[isamp]); a[1] += int(r[iex]); a[2] += float(r[ith]) * int(r[iex])
ts = sum(a[0] for a in agg.values()); te = sum(a[1] for a in agg.values())
print(f"total samples {ts}, warp-instructions {te}")
src = {}
def text(f, n):
    if f not in src:
        for dd in ("isaacgymenvs_b200/csrc", "."):
            p = os.path.join(dd, f)
            if os.path.exists(p):
                src[f] = open(p).read().splitlines(); break
        else:
            src[f] = []
    return src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""
for (f, n), a in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print(f"{100*a[1]/te:5.1f}% exec {100*a[0]/max(ts,1):5.1f}% smp thr {a[2]/max(a[1],1):4.1f}  {f}:{n}  {text(f, n)}")

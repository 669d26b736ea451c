"""Attribute an ncu source-page CSV (SASS view) to CUDA source lines using nvdisasm line info.
usage: ncu_lines.py <report.ncu-rep> <mangled-kernel-substring> [top]"""
import collections, csv, glob, os, re, subprocess, sys, tempfile
rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "ground_fusion_b200", "libgf_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
txt = ""
for f in glob.glob(os.path.join(tmp, "*.cubin")):
    txt += subprocess.run(["nvdisasm", "--print-line-info", f], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
hi = [i for i, r in enumerate(rows) if "Source" in r and "Address" in r][0]
h = rows[hi]
ia, ie, isrc = h.index("Address"), h.index("Instructions Executed"), h.index("Source")
isamp = h.index("Warp Stall Sampling (All Samples)")
data = []
for r in rows[hi + 1:]:
    try:
        data.append((int(r[ia], 16), int(r[ie]), int(r[isamp]), r[isrc]))
    except Exception:
        pass
base = min(d[0] for d in data)
m = re.search(r"\.section\s+(\.text\.[^\s,]*%s[^\s,]*)" % re.escape(kern), txt)
i = m.start()
j = txt.find("\n\t.section\t.text.", i + 20)
seg = txt[i:j if j > 0 else None]
cur, amap = None, {}
for line in seg.splitlines():
    mm = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if mm:
        cur = (mm.group(1).split("/")[-1], int(mm.group(2)))
        continue
    mm = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+[A-Z@!]", line)
    if mm:
        amap[int(mm.group(1), 16)] = cur
ex, sm = collections.Counter(), collections.Counter()
for a, e, s_, _ in data:
    k = amap.get(a - base)
    ex[k] += e; sm[k] += s_
te, ts = sum(ex.values()), sum(sm.values())
print("instructions: %d static, %d executed (warp-level); %d stall samples" % (len(data), te, ts))
lines = {}
for k in set(list(ex) + list(sm)):
    if k and k[0] not in lines:
        for d_ in (os.path.join(root, "ground_fusion_b200", "csrc"), "/usr/local/cuda/include"):
            p = os.path.join(d_, k[0])
            if os.path.exists(p):
                lines[k[0]] = open(p, errors="replace").read().splitlines()
for k, v in sorted(sm.items(), key=lambda kv: -kv[1])[:top]:
    text = ""
    if k and k[0] in lines and k[1] - 1 < len(lines[k[0]]):
        text = lines[k[0]][k[1] - 1].strip()[:90]
    print("%-28s samples %5.1f%%  inst %5.1f%%  %s" % ("%s:%d" % k if k else "?", 100.0 * v / max(ts, 1), 100.0 * ex[k] / max(te, 1), text))

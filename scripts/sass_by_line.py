"""Static instruction count per CUDA source line of one kernel (nvdisasm -g on the cubin inside an object file).
usage: sass_by_line.py OBJECT.o KERNEL [top]"""
import re, collections, subprocess, sys, glob, os, tempfile
obj, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
cub = glob.glob(tmp + "/*.cubin")[0]
dis = subprocess.run(["nvdisasm", "-g", cub], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text." + kern + ":"))
cur, cnt, tot, ops = None, collections.Counter(), 0, collections.Counter()
for l in dis[start + 1:]:
    if l.startswith(".text.") or l.startswith(".section"): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m and cur: cnt[cur] += 1; tot += 1; ops[m.group(1).split(".")[0]] += 1
print("instructions:", tot)
print(" ".join("%s:%d" % kv for kv in ops.most_common(25)))
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = {}
for (f, n), c in cnt.most_common(top):
    if f not in src:
        p = glob.glob(root + "/zpaqsharp_b200/csrc/**/" + f, recursive=True)
        src[f] = open(p[0]).read().split("\n") if p else None
    line = src[f][n - 1].strip()[:100] if src[f] and n <= len(src[f]) else ""
    print("%5d %s:%d  %s" % (c, f, n, line))

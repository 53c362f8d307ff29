"""Turn an .ncu-rep into the summaries kept under profiles/: the raw metric page, and the source page reduced to one row
per CUDA source line (executed instructions, stall samples, top stall reasons) plus a per-kernel stall total.
usage: ncu_summarise.py REPORT.ncu-rep OUT_PREFIX"""
import csv, collections, io, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
open(out + "_raw.csv", "w").write(raw)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True, check=True).stdout
lines, total, H, cur = [], collections.Counter(), None, None
def rows(text):
    """ncu quotes every field but does not escape quotes inside source text (inline asm): split data rows from the right."""
    ncol = None
    for line in text.splitlines():
        if not line.startswith('"'):
            continue
        if line.startswith('"Line No"') or line.startswith('"File Path"') or ncol is None:
            r = next(csv.reader([line]))
            if r and r[0] == "Line No": ncol = len(r)
            yield r
            continue
        body = line[1:-1] if line.endswith('"') else line[1:]
        parts = body.rsplit('","', ncol - 2)
        first = parts[0].split('","', 1)
        yield [first[0], first[1] if len(first) > 1 else ""] + parts[1:]
for r in rows(src):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": H = r; continue
    if H is None or not r[0].isdigit(): continue           # SASS rows have an empty line number
    ie, isamp = H.index("Instructions Executed"), H.index("# Samples")
    st = collections.Counter()
    for i, h in enumerate(H):
        if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit() and int(r[i]): st[h[6:]] += int(r[i])
    n, s = int(r[ie] or 0), int(r[isamp] or 0)
    if n == 0 and s == 0: continue
    total.update(st)
    lines.append((cur, int(r[0]), n, s, " ".join("%s:%d" % kv for kv in st.most_common(3)), r[1].strip()))
lines.sort(key=lambda x: -x[3])
w = csv.writer(open(out + "_by_source_line.csv", "w"))
w.writerow(["file", "line", "instructions_executed", "stall_samples", "top_stalls", "source"])
w.writerows(lines)
tot = sum(total.values())
R = list(csv.reader(io.StringIO(raw)))
for name in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread"):
    if name in R[0]: print(name, R[2][R[0].index(name)], R[1][R[0].index(name)])
print("stall samples", sum(l[3] for l in lines), "(an instruction inlined through several files is listed under each)")
print("stalls:", ", ".join("%s %.1f%%" % (k, 100.0 * v / tot) for k, v in total.most_common(9)))

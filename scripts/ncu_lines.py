"""Per-source-line stall samples from an .ncu-rep (run here, no GPU):  ncu_lines.py rep kernel [top]"""
import csv
import subprocess
import sys

rep, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file = None
hdr = None
agg = {}
reasons = {}
tot = 0
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < len(hdr) or r[2] != "-":
        continue  # only the per-source-line summary rows (Address == "-")
    try:
        s = int(r[hdr.index("# Samples")])
        ins = int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    key = (cur_file, int(r[0]), r[1].strip()[:70])
    a = agg.setdefault(key, [0, 0, {}])
    a[0] += s
    a[1] += ins
    for ci, name in enumerate(hdr):
        if name.startswith("stall_") and "Not Issued" not in name:
            try:
                v = int(r[ci])
            except ValueError:
                continue
            if v:
                a[2][name[6:]] = a[2].get(name[6:], 0) + v
                reasons[name[6:]] = reasons.get(name[6:], 0) + v
    tot += s
print(f"total samples {tot}; stall reasons:", ", ".join(f"{k} {100.0 * v / max(tot, 1):.0f}%" for k, v in sorted(reasons.items(), key=lambda kv: -kv[1])[:8]))
for (f, ln, src), (s, ins, rs) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    top_r = ",".join(f"{k}:{v}" for k, v in sorted(rs.items(), key=lambda kv: -kv[1])[:3])
    print(f"{100.0 * s / max(tot, 1):5.1f}% {ins:9d} inst  {f}:{ln}  {src}  [{top_r}]")

"""Per-source-line stall samples of a kernel from an ncu report (needs -lineinfo and --import-source on):
python tools/ncu_lines.py report.ncu-rep [top_n]"""
import csv, subprocess, sys, collections
rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, hdr, lines = "", None, collections.OrderedDict()
total = 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ix = {}
        for i, h in enumerate(hdr):
            ix.setdefault(h, i)
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or r[0] in ("Function Name",):
        continue
    if r[0] != "":          # a source line with aggregated metrics
        try:
            n = int(r[ix["# Samples"]] or 0)
        except ValueError:
            continue
        key = (cur_file, int(r[0]))
        st = {s: int(r[ix[s]] or 0) for s in stalls}
        if key in lines:
            lines[key][0] += n
            for s in stalls:
                lines[key][2][s] += st[s]
        else:
            lines[key] = [n, r[1].strip(), st, int(r[ix["Instructions Executed"]] or 0)]
        total += n
print(f"total samples {total}")
agg = collections.Counter()
for (f, ln), (n, src, st, ie) in lines.items():
    for s, v in st.items():
        agg[s] += v
print("stall mix: " + ", ".join(f"{s[6:]} {100 * v / max(1, sum(agg.values())):.1f}%" for s, v in agg.most_common(8)))
for (f, ln), (n, src, st, ie) in sorted(lines.items(), key=lambda kv: -kv[1][0])[:top_n]:
    best = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"{100 * n / total:5.1f}% {n:6d}  {f}:{ln:<4d} {best[0][0][6:]:>14s} {best[1][0][6:]:>14s}  {src[:100]}")

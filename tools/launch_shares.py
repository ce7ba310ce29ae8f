"""Per-kernel launch shares of one bench step from an ncu launch list (development tool).
usage: launch_shares.py launches.csv  — prints a markdown table for the LAST step in the list."""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ix = {h: i for i, h in enumerate(hdr)}
L = [(int(r[ix["ID"]]), r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), float(r[ix["Metric Value"]])) for r in rows[1:]]
# the last FASTQ-mode k_stream_ws launch starts the last step
starts = [i for i, (_, n, _) in enumerate(L) if n.startswith("k_stream") and ", 0, " in n or n.endswith(", 0>") or ("k_stream" in n and "false" in n)]
fq = [i for i, (_, n, _) in enumerate(L) if "k_stream" in n and L[i][2] > 5e6]
step = L[fq[-1]:]
agg = collections.OrderedDict()
for _, n, t in step:
    n = n.split("<")[0]
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += t / 1000.0
tot = sum(a[1] for a in agg.values())
print("| kernel | launches per step | total us | share |\n|---|---:|---:|---:|")
for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
    print("| %s | %d | %.1f | %.1f%% |" % (n, a[0], a[1], 100 * a[1] / tot))
print("| **all** | %d | %.1f | 100%% |" % (sum(a[0] for a in agg.values()), tot))

"""top stall-sample instructions of a .ncu-rep source page, with their dominant stall reasons
(development aid): python scripts/ncu_hot.py report.ncu-rep [N]"""
import csv, subprocess, sys
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h, rows = r[1], r[2:]
ix = {k: i for i, k in enumerate(h)}
keys = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(x[2]) for x in rows)
print("total samples", tot, "instructions", len(rows))
agg = {k: sum(int(x[ix[k]]) for x in rows) for k in keys}
print({k[6:]: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v > tot * 0.02})
for i in sorted(sorted(range(len(rows)), key=lambda i: -int(rows[i][2]))[:n]):
    x = rows[i]
    st = {k[6:]: int(x[ix[k]]) for k in keys if int(x[ix[k]]) > int(x[2]) * 0.2}
    print(i, x[2], x[5], x[1][:72], st)

"""text summary of a .ncu-rep (details page): `python scripts/ncu_details.py report.ncu-rep > profiles/x.txt`
(section, metric, unit, value per line - the form the files under profiles/ are kept in)"""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "details", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]; ix = {k: i for i, k in enumerate(h)}
name = None
for r in rows[1:]:
    if len(r) <= ix["Metric Value"] or not r[ix["Metric Name"]]:
        continue                  # rule rows / truncated rows carry no metric
    if r[ix["Kernel Name"]] != name:
        name = r[ix["Kernel Name"]]
        print("kernel:", name)
    print(f'{r[ix["Section Name"]][:28]:28s} {r[ix["Metric Name"]][:44]:44s} {r[ix["Metric Unit"]]:14s} {r[ix["Metric Value"]]}')

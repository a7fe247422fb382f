"""DRAM bytes and executed FP64 flops per launch of the kernels in .ncu-rep files (ncu --set full), in the form
profiles/traffic.json keeps them:  python scripts/ncu_traffic.py report.ncu-rep [...]"""
import csv, json, subprocess, sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
out = {}
for rep in sys.argv[1:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(txt.splitlines()))
    h, u = r[0], r[1]
    for v in r[2:]:
        d = {k: (x, un) for k, un, x in zip(h, u, v)}
        name = d["Kernel Name"][0].split("(")[0].replace("void ", "")
        byts = sum(float(d[k][0]) * UNIT[d[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        cyc = float(d["smsp__cycles_elapsed.max"][0]) if "smsp__cycles_elapsed.max" in d else float(d["sm__cycles_elapsed.max"][0])
        per = {op: float(d[f"smsp__sass_thread_inst_executed_op_{op}_pred_on.sum.per_cycle_elapsed"][0]) for op in ("dadd", "dmul", "dfma")}
        flops = (per["dadd"] + per["dmul"] + 2 * per["dfma"]) * cyc
        out[name] = {"report": rep.split("/")[-1], "dram_bytes_per_launch": byts, "fp64_flops_per_launch": flops,
                     "duration_us_under_ncu": float(d["gpu__time_duration.sum"][0]) * (1e3 if d["gpu__time_duration.sum"][1] == "ms" else 1.0)}
print(json.dumps(out, indent=1))

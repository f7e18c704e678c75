"""profiles/dram_traffic.json from an ncu metrics pass over one resident pass of the bench workload.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
        --log-file gpurun_out/r02_dram_bytes.csv python tools/k1_prof.py cfg2_150bp 1000000
    python tools/ncu_traffic.py gpurun_out/r02_dram_bytes.csv 4

tools/k1_prof.py runs the resident pass four times (one inside bsw_resident_create, two warm, one reported); the bytes are summed per kernel family over
all launches and divided by the number of passes (second argument).  bench.py reads the file for `roofline.traffic`."""
import csv, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1]
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
rows = [l for l in open(path) if l.startswith('"')]
fam = {"k0": 0.0, "k1": 0.0}
launches = {"k0": set(), "k1": set()}
names = {}
for r in csv.DictReader(rows):
    if r["Metric Name"] not in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        continue
    v = float(r["Metric Value"].replace(",", ""))
    v *= {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[r["Metric Unit"].lower()]
    k = "k1" if "k1_extend" in r["Kernel Name"] else "k0"
    fam[k] += v
    launches[k].add(r["ID"])
    short = r["Kernel Name"].split("(")[0]
    names[short] = names.get(short, 0) + 1
try:
    commit = subprocess.check_output(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], text=True).strip()
except Exception:
    commit = "unknown"
out = {"k0": round(fam["k0"] / passes), "k1": round(fam["k1"] / passes),
       "launches_per_pass": {k: len(v) / passes for k, v in launches.items()},
       "kernels": {k: v // 2 for k, v in names.items()},
       "capture": f"ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, tools/k1_prof.py cfg2_150bp 1000000, {passes} passes averaged, "
                  f"{os.path.basename(path)}, source tree after commit {commit}",
       "note": "k0 = everything that is not K1 (device scheduler + gather)"}
json.dump(out, open(os.path.join(ROOT, "profiles", "dram_traffic.json"), "w"), indent=1)
print(json.dumps(out))

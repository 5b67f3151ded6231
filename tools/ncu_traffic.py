"""DRAM traffic of a kernel from an `ncu --set full` report -> profiles/traffic_<name>.json (what bench.py quotes as roofline.traffic).

    python tools/ncu_traffic.py gpurun_out/X.ncu-rep <name> "<what was captured>"

Reads `ncu -i X.ncu-rep --page raw --csv` (works without a GPU) and sums dram__bytes_read.sum + dram__bytes_write.sum of the
first kernel in the report."""
import csv, io, json, os, subprocess, sys

rep, name, what = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, first = rows[0], rows[1], rows[2]


def col(metric):
    i = hdr.index(metric)
    v = float(first[i].replace(",", ""))
    u = units[i].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "tbyte": 1e12}.get(u, 1)


rd, wr = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
dur_i = hdr.index("gpu__time_duration.sum")
out = {"kernel": first[hdr.index("Kernel Name")], "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
       "duration_under_ncu": first[dur_i] + " " + units[dur_i],
       "source": f"ncu --set full --clock-control none, {os.path.basename(rep)} summarised in profiles/ ({what})"}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", f"traffic_{name}.json")
json.dump(out, open(path, "w"), indent=1)
print(path, out)

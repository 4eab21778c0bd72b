"""update_traffic.py <file.ncu-rep> <launch index> <key in profiles/traffic.json>: DRAM bytes (read + write) of one
profiled launch of an `ncu --set full` capture, written into profiles/traffic.json (bench.py's roofline.traffic)."""
import csv, json, subprocess, sys

rep, idx, key = sys.argv[1], int(sys.argv[2]), sys.argv[3]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
total = 0.0
for name in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
    i = hdr.index(name)
    total += float(data[idx][i].replace(",", "")) * scale[units[i]]
path = "profiles/traffic.json"
t = json.load(open(path))
t[key] = int(round(total))
json.dump(t, open(path, "w"), indent=2)
print(key, data[idx][hdr.index("Kernel Name")], int(round(total)))

"""Write profiles/r02_decode_traffic.json (what bench.py's roofline.traffic scales from) out of an ncu --set full capture
of one zxc_decode_kernel launch over N MiB of the bench corpus:  python profiles/update_traffic.py REP.ncu-rep MiB NAME"""
import csv, json, os, subprocess, sys
rep, mib, name = sys.argv[1], int(sys.argv[2]), sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u, v = rows[0], rows[1], rows[-1]
def get(metric):
    i = h.index(metric)
    x = float(v[i].replace(",", ""))
    return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u[i]]
out = {"kernel": v[h.index("Kernel Name")] if "Kernel Name" in h else "zxc_decode_kernel",
       "dram_bytes_read": int(get("dram__bytes_read.sum")), "dram_bytes_write": int(get("dram__bytes_write.sum")),
       "decoded_bytes": mib << 20,
       "compressed_bytes_note": "%d MiB of the bench corpus at level 3 / 64 KiB blocks (ratio 0.405)" % mib,
       "source": "profiles/%s (ncu --set full --clock-control none, one launch over %d MiB)" % (name, mib)}
root = os.path.dirname(os.path.abspath(__file__))
json.dump(out, open(os.path.join(root, "r02_decode_traffic.json"), "w"), indent=1)
print(json.dumps(out))

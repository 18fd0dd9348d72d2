"""profiles/r01_demb_ncu_full_summary.csv -> profiles/kernel_traffic.json (dram read + write bytes per launch, unit-aware)."""
import csv, json, re, sys
src = sys.argv[1] if len(sys.argv) > 1 else "profiles/r01_demb_ncu_full_summary.csv"
dst = sys.argv[2] if len(sys.argv) > 2 else "profiles/kernel_traffic.json"
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
rows = list(csv.reader(open(src)))
h = rows[0]
def col(prefix):
    i = [k for k, c in enumerate(h) if c.startswith(prefix)][0]
    return i, UNIT[re.search(r"\[(\w+)\]", h[i]).group(1)]
(ri, ru), (wi, wu) = col("dram__bytes_read.sum"), col("dram__bytes_write.sum")
t = {}
for r in rows[1:]:
    name = re.sub(r"^(void )?(<unnamed>::)?", "", r[0]); name = re.sub(r"<.*$", "", name)
    t[name] = int(float(r[ri]) * ru + float(r[wi]) * wu)
t["forward_seq_kernel"] = t.get("forward_seq_tma_kernel")
t["_source"] = f"{src} (ncu --set full of `bench.py --ncu`, dram__bytes_read.sum + dram__bytes_write.sum per launch)"
json.dump(t, open(dst, "w"), indent=1)
print(t)

"""DRAM traffic per igemm launch from one `ncu --set full` report -> profiles/igemm_traffic.json (bench.py's roofline.traffic).
    python tools/ncu_traffic.py gpurun_out/r02_full.ncu-rep "source note" > profiles/igemm_traffic.json"""
import csv
import json
import subprocess
import sys

rep = sys.argv[1]
note = sys.argv[2] if len(sys.argv) > 2 else rep
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")


def to_bytes(v, unit):
    v = float(v.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


per = [to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in rows[2:] if "igemm_kernel" in r[ik]]
print(json.dumps({"dram_bytes_per_launch": int(sum(per) / max(1, len(per))),
                  "source": f"ncu --set full ({note}): dram__bytes_read.sum + dram__bytes_write.sum, mean over the first {len(per)} "
                            "igemm launches of one DDIM step",
                  "per_launch_MB": [round(b / 1e6, 3) for b in per]}, indent=1))

"""Hot source lines of one kernel from an ncu report (stall samples per CUDA source line).
    python tools/ncu_source_hot.py <rep> <kernel-regex> [launch-skip] [file-substr]
"""
import csv, subprocess, sys

def num(x):
    try: return int(float(x.replace(",", "")))
    except Exception: return 0

rep, kre = sys.argv[1], sys.argv[2]
skip = sys.argv[3] if len(sys.argv) > 3 else "0"
want = sys.argv[4] if len(sys.argv) > 4 else ".cu"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre, "--launch-skip", skip,
                      "--launch-count", "1", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == "File Path":
        path = rows[i][1]; hdr = rows[i + 2]; ix = {h: j for j, h in enumerate(hdr)}
        j = i + 3; body = []
        while j < len(rows) and not (rows[j] and rows[j][0] == "File Path"):
            body.append(rows[j]); j += 1
        if want in path:
            stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
            per = []; tot = 0
            for r in body:
                if r and r[0].isdigit():
                    s = num(r[4]); tot += s
                    per.append((int(r[0]), r[1].strip(), s, num(r[7]), {h: num(r[ix[h]]) for h in stalls}))
            print(f"# {path}: {rows[i+1][1][:100]}  total samples {tot}")
            for ln, src, s, ie, d in sorted(per, key=lambda x: -x[2])[:int(sys.argv[5]) if len(sys.argv) > 5 else 30]:
                top = [(k[6:], v) for k, v in sorted(d.items(), key=lambda kv: -kv[1])[:3] if v]
                print(f"{ln:4d} {100.0*s/max(tot,1):5.1f}% inst={ie:9d} {src[:78]:78s} {top}")
        i = j
    else:
        i += 1

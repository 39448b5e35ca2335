"""Per-kernel SASS evidence of the Blackwell-native paths in libvdb200.so (B200_PROFILING.md: tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA loads/stores -> UTMALDG/UTMASTG, legacy tensor path -> HMMA).
    python tools/sass_summary.py [lib] > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "versatile-diffusion_b200", "vdb200", "libvdb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
pats = collections.OrderedDict([("UTC*MMA", r"\bUTC\w*MMA"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTMALDG", r"\bUTMALDG"),
                                ("UTMASTG", r"\bUTMASTG"), ("MUFU", r"\bMUFU"), ("HMMA", r"\bHMMA"), ("instr", r"^\s+/\*[0-9a-f]{4}\*/")])
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur:
        for k, p in pats.items():
            if re.search(p, line):
                counts[cur][k] += 1
print(f"# {os.path.basename(lib)}: {len(counts)} kernels; SASS instruction counts per kernel (cuobjdump -sass)")
print(f"{'kernel':110s} " + " ".join(f"{k:>8s}" for k in pats))
tot = collections.Counter()
fam = collections.OrderedDict()
for name, c in counts.items():
    d = demangle(name)
    d = re.sub(r"\(vdb::\w+Params\)|\(.*\)$", "", d).replace("void vdb::", "").replace("(int)", "")
    tot.update(c)
    base = d.split("<")[0]
    f = fam.setdefault(base, [0, collections.Counter()])
    f[0] += 1
    f[1].update(c)
    if any(c[k] for k in ("UTC*MMA", "LDTM", "STTM", "UTMALDG", "UTMASTG")):
        print(f"{d[:110]:110s} " + " ".join(f"{c[k]:8d}" for k in pats))
print("\n# by kernel family (instantiations summed)")
for base, (n, c) in fam.items():
    print(f"{(base + ' x' + str(n))[:110]:110s} " + " ".join(f"{c[k]:8d}" for k in pats))
print(f"\n{'TOTAL':110s} " + " ".join(f"{tot[k]:8d}" for k in pats))

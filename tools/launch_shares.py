"""ncu launch list (csv of gpu__time_duration.sum, tools/round_check.sh) -> per-kernel share table.
The profiled range is one eager DDIM step followed by the VAE decode (tools/profile_step.py --decode);
the first `permute_f32_kernel` after the UNet step marks where the decode starts.

    python tools/launch_shares.py gpurun_out/launches_v7.csv > profiles/r01_launch_shares_v7.txt
"""
import csv
import re
import sys
from collections import OrderedDict


def short(name):
    name = name.replace("void ", "").replace("vdb::", "")
    m = re.match(r"([A-Za-z0-9_:]+)(<[^(]*>)?\(", name)
    if not m:
        return name[:48]
    base, targs = m.group(1), m.group(2) or ""
    if base.startswith("at::"):
        return base
    if "igemm" in base or "attention" in base:
        return base.replace("_kernel", "") + targs
    return base


def table(rows, title):
    tot = sum(ns for _, ns in rows)
    agg = OrderedDict()
    for k, ns in rows:
        a = agg.setdefault(k, [0, 0])
        a[0] += 1
        a[1] += ns
    print(f"{title}: {len(rows)} launches, {tot / 1e3:.0f} us (ncu gpu__time_duration.sum, --clock-control none, serialised)")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"   {k:44s} n={n:4d} {ns / 1e3:9.1f} us {100.0 * ns / tot:5.1f}%  avg {ns / 1e3 / n:7.1f} us")


def main(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        rows.append((short(r["Kernel Name"]), float(r["Metric Value"].replace(",", ""))))
    # the sampler's final NHWC->NCHW permute of the latent ends the step; everything after is the decode
    cut = len(rows)
    for i, (k, _) in enumerate(rows):
        if k == "ddim_cfg_step_kernel":
            cut = i + 2          # + the step-counter decrement
    step, rest = rows[:cut], rows[cut:]
    table(step, "ONE DDIM STEP (B=8, 64x64 latent, CFG) incl. per-sample setup kernels")
    if rest:
        table(rest, "AFTER THE STEP (latent permute + VAE DECODE of 4 images 512x512)")


if __name__ == "__main__":
    main(sys.argv[1])

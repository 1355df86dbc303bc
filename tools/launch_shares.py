#!/usr/bin/env python
"""Condenses an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals and shares.
Usage: python tools/launch_shares.py gpurun_out/launches.csv "<the command that was profiled>" > profiles/rNN_launch_shares.txt"""
import collections
import csv
import sys

MICRO = ("red_peak_kernel", "fma_peak_kernel", "dfma_peak_kernel", "tile_rmw_peak_kernel")


def main(path, cmd):
    hdr, tot = None, collections.OrderedDict()
    for r in csv.reader(open(path)):
        if hdr is None:
            if "Kernel Name" in r:
                hdr, ki, vi = r, r.index("Kernel Name"), r.index("Metric Value")
            continue
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("void ", "")
        n, t = tot.get(name, (0, 0.0))
        tot[name] = (n + 1, t + float(r[vi].replace(",", "")) / 1e6)
    print(cmd)
    print("(per-launch times are cold-cache and serialised: compare SHARES; the *_peak kernels are the roofline "
          "microbenchmarks, not part of a step)")
    all_ms = sum(t for _, t in tot.values())
    for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:50s} n={n:3d} total {t:9.3f} ms  share {100 * t / all_ms:5.1f}%")
    step = {k: v for k, v in tot.items() if not k.startswith(MICRO)}
    step_ms = sum(t for _, t in step.values())
    print("\nshares within the steps (microbenchmarks excluded):")
    for k, (n, t) in sorted(step.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:50s} share {100 * t / step_ms:5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")

#!/usr/bin/env python
"""Basic-block view of an `ncu --set full --import-source on` capture: groups the SASS of the kernel into runs of
instructions with the same execution count (= basic blocks as executed) and prints, per block, its share of the executed
warp instructions and of the stall samples, its instruction mix and its top stall reasons.
    python tools/ncu_blocks.py gpurun_out/prof_trace.ncu-rep [min_share_percent] > profiles/r02_trace_blocks.txt"""
import collections
import csv
import io
import subprocess
import sys


def main(path, min_share=0.2):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    kernel = rows[0][1] if rows and rows[0][0] == "Kernel Name" else "?"
    hdr = rows[1]
    col = {n: i for i, n in enumerate(hdr)}
    stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
    blocks, cur = [], None
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        execs = int(float(r[col["Instructions Executed"]] or 0))
        op = r[col["Source"]].split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "?")
        op = op.rstrip(";")
        branchy = op.split(".")[0] in ("BRA", "EXIT", "BSYNC", "RET", "CALL", "BRX", "WARPSYNC")
        if cur is None or cur["execs"] != execs:
            cur = {"execs": execs, "n": 0, "samples": 0, "ops": collections.Counter(), "stalls": collections.Counter(),
                   "wavefronts": 0}
            blocks.append(cur)
        cur["n"] += 1
        cur["samples"] += int(float(r[col["# Samples"]] or 0))
        cur["ops"][op] += 1
        cur["wavefronts"] += int(float(r[col["L1 Wavefronts Shared"]] or 0))
        for s in stall_cols:
            cur["stalls"][s[6:]] += int(float(r[col[s]] or 0))
        if branchy:
            cur = None
    tot_i = sum(b["n"] * b["execs"] for b in blocks) or 1
    tot_s = sum(b["samples"] for b in blocks) or 1
    print(kernel)
    print(f"ncu --set full --import-source on, --page source: basic blocks by share of executed warp instructions "
          f"(total {tot_i}, {tot_s} stall samples)\n")
    for b in sorted(blocks, key=lambda b: -b["n"] * b["execs"]):
        share = 100.0 * b["n"] * b["execs"] / tot_i
        if share < min_share:
            continue
        ops = ", ".join(f"{o} x{c}" for o, c in b["ops"].most_common(6))
        st = ", ".join(f"{s} {100 * c // max(1, b['samples'])}%" for s, c in b["stalls"].most_common(4))
        wf = f" | {b['wavefronts']} shared-memory wavefronts" if b["wavefronts"] else ""
        print(f"{b['n']:4d} instructions x {b['execs']:9d} executions: {share:5.1f}% of instructions, "
              f"{100.0 * b['samples'] / tot_s:5.1f}% of samples | {ops} | stalls: {st}{wf}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.2)

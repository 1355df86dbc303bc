#!/usr/bin/env python
"""profiles/r02_traffic.json: DRAM traffic of the two hot kernels from their `ncu --set full` captures, per ray and per
segment of the capture, so that bench.py can scale it to its own launch sizes and name its source instead of carrying
typed-in constants.
    python tools/ncu_traffic.py gpurun_out/prof_trace.ncu-rep gpurun_out/prof_trace.log \\
                                gpurun_out/prof_tile_raster.ncu-rep gpurun_out/prof_tile_raster.log
The .log files are the bench lines of the two capture runs (bench.py --rays-per-gpu N --steps 1 --no-extras): rays and
segments per launch come from there."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def metrics(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))

    def bytes_of(name):
        v = float(d[name].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[name]]
    return {"kernel": d.get("Kernel Name"), "dram_read": bytes_of("dram__bytes_read.sum"),
            "dram_write": bytes_of("dram__bytes_write.sum"),
            "time_ms": float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "msecond": 1.0,
                                                                                "usecond": 1e-3, "nsecond": 1e-6}[u["gpu__time_duration.sum"]]}


def bench_line(log):
    for line in open(log):
        if line.startswith("{"):
            return json.loads(line)
    raise SystemExit(f"no bench line in {log}")


def main(trace_rep, trace_log, raster_rep, raster_log, tag="r02"):
    t, b = metrics(trace_rep), bench_line(trace_log)
    rays = b["config"]["rays_per_gpu"]
    out = {"trace": {"file": f"profiles/{tag}_trace_full.txt", "kernel": t["kernel"], "rays": rays,
                     "segments": b["segments_per_ray"] * rays, "dram_bytes": t["dram_read"] + t["dram_write"],
                     "dram_bytes_read": t["dram_read"], "dram_bytes_write": t["dram_write"],
                     "dram_bytes_per_ray": (t["dram_read"] + t["dram_write"]) / rays, "kernel_ms": t["time_ms"]}}
    r, b = metrics(raster_rep), bench_line(raster_log)
    rays = b["config"]["rays_per_gpu"]
    segs = b["segments_per_ray"] * rays
    out["tile_raster"] = {"file": f"profiles/{tag}_tile_raster_full.txt", "kernel": r["kernel"], "rays": rays, "segments": segs,
                          "fragments": b["pixel_updates_per_s_in_kernel"] * b["phase_ms_per_step"]["accumulate"] * 1e-3,
                          "dram_bytes": r["dram_read"] + r["dram_write"], "dram_bytes_read": r["dram_read"],
                          "dram_bytes_write": r["dram_write"],
                          "dram_bytes_per_segment": (r["dram_read"] + r["dram_write"]) / segs, "kernel_ms": r["time_ms"]}
    out["how"] = ("ncu --set full --clock-control none --import-source on -k regex:<kernel> -s 1 -c 1 python bench.py "
                  "--rays-per-gpu 2000000 --steps 1 --warmup 3 --no-cpu-baseline --no-extras (tools/gpu_r02.sh proftrace / "
                  "profraster); dram__bytes_read.sum + dram__bytes_write.sum of that one launch")
    path = os.path.join(ROOT, "profiles", f"{tag}_traffic.json")
    json.dump(out, open(path, "w"), indent=1)
    print(path)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main(*sys.argv[1:])

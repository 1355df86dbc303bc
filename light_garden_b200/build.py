"""Builds the CUDA library in-tree: light_garden_b200/_lib/liblight_garden_b200.so.

nvcc cross-compiles sm_100a without a GPU.  Flags that matter for parity:
  -fmad=false                the only fused multiply-adds are the explicit ones of ORACLE.md
  -Xcompiler -ffp-contract=off   same for the host-side lowering code
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
OBJ_DIR = os.path.join(OUT_DIR, "obj")
LIB = os.path.join(OUT_DIR, "liblight_garden_b200.so")
SOURCES = ["lg_capi.cu", "lg_trace_f32.cu", "lg_trace_f64.cu", "lg_trace_grid.cu", "lg_trace_f32_dup.cu", "lg_trace_f64_large.cu"]
HEADERS = ["lg_geom.cuh", "lg_nearest.cuh", "lg_trace.cuh", "lg_accum.cuh", "lg_tiles.cuh", "lg_nested.cuh", "lg_reduce.cuh", "lg_bench.cuh", "lg_srgb.h", "lg_scene.h", "lg_tables.h", "../../include/light_garden_b200.h"]

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CCBIN = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
FLAGS = [
    "-ccbin", CCBIN,
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall",
]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False, ptxas_v=False):
    os.makedirs(OBJ_DIR, exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS] + [os.path.abspath(__file__)]
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ_DIR, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if ptxas_v else []) + ["-c", s, "-o", o]
            jobs.append(cmd)

    def run(cmd):
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    logs = []
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        logs = list(ex.map(run, jobs))
    objs = [os.path.join(OBJ_DIR, s.replace(".cu", ".o")) for s in SOURCES]
    if jobs or force or _stale(LIB, objs):
        cmd = [NVCC, "-ccbin", CCBIN, "-shared", "-o", LIB] + objs + ["-ldl"]
        run(cmd)
    if ptxas_v:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(verbose=True, force="--force" in sys.argv, ptxas_v="--ptxas" in sys.argv))

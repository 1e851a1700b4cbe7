"""Build libomnissm.so (sm_100a) in-tree with nvcc.  No torch, no cmake.

    python -m omnimamba_b200.build [--force] [--verbose]

Each csrc/*.cu is compiled to an object in parallel and linked into
omnimamba_b200/lib/libomnissm.so.  Objects are rebuilt only when the source (or a header) is newer.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build")
LIB = os.path.join(LIBDIR, "libomnissm.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
              "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libomnissm.so cannot be built")


def _newest_header_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for dp, _, fns in os.walk(root):
            for fn in fns:
                if fn.endswith((".h", ".cuh", ".hpp")):
                    m = max(m, os.path.getmtime(os.path.join(dp, fn)))
    return m


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdr_m = _newest_header_mtime()
    jobs = []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJDIR, s[:-3] + ".o")
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_m)
        jobs.append((src, obj, stale))

    def compile_one(job):
        src, obj, stale = job
        if not stale:
            return src, 0, ""
        # OMNI_NVCC_EXTRA: extra compile flags for experiment builds (e.g. -DOMNI_TC_VARIANTS)
        cmd = [nvcc, *ARCH, *NVCC_FLAGS, *os.environ.get("OMNI_NVCC_EXTRA", "").split(), "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r.returncode, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(8, len(jobs) or 1)) as ex:
        results = list(ex.map(compile_one, jobs))
    log = []
    for src, rc, out in results:
        if out:
            log.append(f"== {os.path.basename(src)}\n{out}")
        if rc != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    if log:
        with open(os.path.join(OBJDIR, "ptxas.log"), "w") as f:
            f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    need_link = force or not os.path.exists(LIB) or any(j[2] for j in jobs) or \
        any(os.path.getmtime(j[1]) > os.path.getmtime(LIB) for j in jobs)
    if need_link:
        # the driver API (cuTensorMapEncodeTiled) is resolved at run time through cudaGetDriverEntryPoint,
        # so only the (static) runtime is linked
        cmd = [nvcc, *ARCH, "-shared", "-o", LIB, *[j[1] for j in jobs], "-Xcompiler", "-fPIC", "-cudart", "static"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link of libomnissm.so failed")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)

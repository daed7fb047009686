"""Builds hrm_b200/lib/libhrmgpu.so (sm_100a only) with nvcc.  In-tree, so the .so travels with the repo snapshot.

    python -m hrm_b200.build [--force] [--verbose]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libhrmgpu.so")
HOST_DIR = os.path.join(HERE, "host")
HOST_LIB = os.path.join(LIBDIR, "libhrmhost.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off,-Wall", "--expt-relaxed-constexpr"]
# per-file additions: tfe.cu mirrors host arithmetic operation for operation (no multiply-add contraction)
FILE_FLAGS = {"tfe.cu": ["-fmad=false"]}


def sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "hrmgpu.h"))
    jobs = []
    objs = []
    for s in sources():
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            cmd = [NVCC] + ARCH + FLAGS + FILE_FLAGS.get(s, []) + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0 or verbose:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    if force or jobs or _stale(LIB, objs):
        run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs)
    # C++ host layer (planner API mirror) on top of the C ABI; no CUDA in it, linked against libhrmgpu.so
    host_src = [os.path.join(HOST_DIR, f) for f in sorted(os.listdir(HOST_DIR)) if f.endswith(".cpp")]
    host_dep = host_src + [os.path.join(HOST_DIR, f) for f in os.listdir(HOST_DIR) if f.endswith(".hpp")] + [LIB, headers[-1]]
    if force or _stale(HOST_LIB, host_dep):
        run([os.environ.get("HRM_CXX", "g++"), "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-Wall", "-shared", "-o", HOST_LIB] + host_src +
            ["-L" + LIBDIR, "-lhrmgpu", "-Wl,-rpath,$ORIGIN"])
    # command-line twins of the reference's benchmark drivers
    apps = os.path.join(HOST_DIR, "apps")
    for f in sorted(os.listdir(apps)):
        if f.endswith(".cpp"):
            exe = os.path.join(LIBDIR, f[:-4])
            if force or _stale(exe, [os.path.join(apps, f)] + host_dep):
                run([os.environ.get("HRM_CXX", "g++"), "-O2", "-std=c++17", "-ffp-contract=off", "-Wall", "-o", exe, os.path.join(apps, f),
                     "-I" + os.path.join(HERE, "..", "include"), "-L" + LIBDIR, "-lhrmgpu", "-Wl,-rpath,$ORIGIN"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))

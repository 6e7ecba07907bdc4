"""Build libpecs_b200.so in-tree (pecs_b200/lib) with nvcc for sm_100a.

    python -m pecs_b200.build [--force] [--jobs N]

Host C++ (pecs_b200/csrc/host, error.cpp) is compiled with g++ -fopenmp, CUDA sources with
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo; the link step produces one shared library that carries the
whole C ABI of include/pecs_b200.h and include/pecs_b200_host.h.  nvcc cross-compiles without a GPU.
"""
import argparse
import concurrent.futures as cf
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "lib", "libpecs_b200.so")
NVCC = os.environ.get("PECS_NVCC", shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc")
HOST_CXX = os.environ.get("PECS_HOST_CXX", "/usr/bin/g++")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ARCH + ["-O3", "-lineinfo", "-std=c++17", "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC,-fopenmp,-O3",
                     "-Xptxas", "-v", "--expt-relaxed-constexpr"]
CXX_FLAGS = ["-O3", "-march=x86-64-v3", "-fopenmp", "-fPIC", "-std=c++17", "-Wall", "-I/usr/local/cuda/include"]


def sources():
    cpp = sorted(glob.glob(os.path.join(CSRC, "host", "*.cpp"))) + [os.path.join(CSRC, "error.cpp")]
    cu = sorted(glob.glob(os.path.join(CSRC, "cuda", "*.cu")))
    return cpp, cu


# TEST library (csrc/selftest): CPU checkers of the device formulas and the setup tables, linked AGAINST the product
# library, never into it (include/pecs_b200_selftest.h)
SELFTEST_LIB = os.path.join(HERE, "lib", "libpecs_b200_selftest.so")


def selftest_sources():
    return sorted(glob.glob(os.path.join(CSRC, "selftest", "*.cpp")))


def headers():
    pats = ["*.hpp", "*.cuh", "host/*.hpp", "cuda/*.cuh", "selftest/*.hpp", "../../include/*.h"]
    return [h for p in pats for h in glob.glob(os.path.join(CSRC, p))]


# Experiment builds (never loaded by the product): "nc" = the round-1 non-coherent loads of step-varying vectors, the
# A side of scripts/race_repro.py.  They get their own object directory and library name.
VARIANTS = {"nc": ["-DPECS_B200_NC_STEP_VECTORS=1"], "trace": ["-DPECS_B200_TRACE=1"]}


def variant_paths(variant):
    if not variant:
        return OBJ, LIB
    return OBJ + "_" + variant, os.path.join(HERE, "lib", f"libpecs_b200_{variant}.so")


def compile_one(src, force, newest_header, obj_dir=OBJ, extra=()):
    obj = os.path.join(obj_dir, os.path.relpath(src, CSRC).replace(os.sep, "_") + ".o")
    if not force and os.path.exists(obj) and os.path.getmtime(obj) >= max(os.path.getmtime(src), newest_header):
        return obj, ""
    if src.endswith(".cu"):
        cmd = [NVCC] + NVCC_FLAGS + list(extra) + ["-c", src, "-o", obj]
    else:
        cmd = [HOST_CXX] + CXX_FLAGS + list(extra) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"compile failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return obj, r.stderr


def build(force=False, jobs=None, verbose=False, variant=None):
    OBJ, LIB = variant_paths(variant)
    extra = VARIANTS[variant] if variant else []
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cpp, cu = sources()
    newest_header = max(os.path.getmtime(h) for h in headers())
    with cf.ThreadPoolExecutor(max_workers=jobs or os.cpu_count()) as ex:
        results = list(ex.map(lambda s: compile_one(s, force, newest_header, OBJ, extra), cpp + cu))
    objs = [o for o, _ in results]
    log = "\n".join(e for _, e in results if e)
    if verbose and log:
        print(log)
    with open(os.path.join(OBJ, "ptxas.log"), "a") as f:
        f.write(log)
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(o) for o in objs):
        cmd = [NVCC] + ARCH + ["-shared", "-ccbin", HOST_CXX, "-Xcompiler", "-fopenmp", "-o", LIB] + objs + \
              ["-lcublas", "-lcusolver", "-lgomp"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    if not variant:
        tobjs = [compile_one(src, force, newest_header, OBJ)[0] for src in selftest_sources()]
        if tobjs and (force or not os.path.exists(SELFTEST_LIB) or
                      os.path.getmtime(SELFTEST_LIB) < max(os.path.getmtime(o) for o in tobjs + [LIB])):
            cmd = [HOST_CXX, "-shared", "-fopenmp", "-o", SELFTEST_LIB] + tobjs + \
                  ["-L" + os.path.dirname(LIB), "-lpecs_b200", "-Wl,-rpath,$ORIGIN"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"link failed: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--jobs", type=int, default=None)
    ap.add_argument("-v", "--verbose", action="store_true")
    ap.add_argument("--variant", choices=sorted(VARIANTS), default=None)
    a = ap.parse_args()
    print(build(a.force, a.jobs, a.verbose, a.variant))

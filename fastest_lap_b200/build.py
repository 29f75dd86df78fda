"""Builds fastest_lap_b200/libfastestlap_b200.so in-tree: generates the node programs (codegen/generate.py), compiles one
translation unit per variant in parallel with nvcc for sm_100a, and links them with the C-ABI layer.

    python -m fastest_lap_b200.build [--force]
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.environ.get("FL_OBJ", os.path.join(CSRC, "build"))
LIB = os.environ.get("FL_OUT", os.path.join(HERE, "libfastestlap_b200.so"))
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v"] + os.environ.get("FL_EXTRA_FLAGS", "").split()


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd, log):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout[-4000:]))


def build(force=False, verbose=False):
    from .codegen import generate
    os.makedirs(OBJ, exist_ok=True)
    names = list(generate.VARIANTS)
    codegen_src = [os.path.join(HERE, "codegen", f) for f in ("expr.py", "emit.py", "models.py", "nlpnode.py", "generate.py")]
    for n in names:
        hdr = os.path.join(CSRC, "generated", "gen_%s.cuh" % n)
        if force or not _newer(hdr, codegen_src):
            generate.generate(n, generate.VARIANTS[n])
            os.utime(hdr, None)
    jobs = []
    kern = os.path.join(CSRC, "nlp_kernels.cuh")
    for n in names:
        obj = os.path.join(OBJ, "variant_%s.o" % n)
        src = [os.path.join(CSRC, "variant.cu"), kern, os.path.join(CSRC, "generated", "gen_%s.cuh" % n)]
        if force or not _newer(obj, src):
            jobs.append(([NVCC] + ARCH + FLAGS + ["-DFL_VARIANT=" + n, "-c", src[0], "-o", obj], os.path.join(OBJ, "variant_%s.log" % n)))
    extra_src = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC)) if f.endswith(".cu") and f != "variant.cu"]
    extra_obj = []
    for s in extra_src:
        obj = os.path.join(OBJ, os.path.basename(s)[:-3] + ".o")
        extra_obj.append(obj)
        deps = [s, kern, os.path.join(HERE, "..", "include", "fastestlap_b200.h")] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
        if force or not _newer(obj, deps):
            jobs.append(([NVCC] + ARCH + FLAGS + ["-c", s, "-o", obj], obj[:-2] + ".log"))
    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            list(ex.map(lambda j: _run(*j), jobs))
    objs = [os.path.join(OBJ, "variant_%s.o" % n) for n in names] + extra_obj
    if force or jobs or not _newer(LIB, objs):
        _run([NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart"], os.path.join(OBJ, "link.log"))
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)

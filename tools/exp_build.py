"""Experiment builds of one variant's derivative kernel: python tools/exp_build.py NAME NP_ALL DBLOCK MINB [extra nvcc flags]
Environment: FL_SEGMENT (fenced segment length), EXP_VARIANT.  Writes exp/NAME/lib.so = the experimental unit linked with
stubs for the other variants."""
import os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
name, np_all, dblock, minb = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
extra = sys.argv[5:]
variant = os.environ.get("EXP_VARIANT", "f1_direct")
os.environ["FL_NP_ALL"] = str(np_all)
os.environ["FL_NP_HESS"] = os.environ.get("FL_NP_HESS", str(np_all))
from fastest_lap_b200.codegen import generate as G
d = os.path.join(ROOT, "exp", name)
os.makedirs(os.path.join(d, "generated"), exist_ok=True)
G.OUT_DIR = os.path.join(d, "generated")
print(G.generate(variant, G.VARIANTS[variant]))
csrc = os.path.join(ROOT, "fastest_lap_b200", "csrc")
for f in ("variant.cu", "nlp_kernels.cuh", "fl_math.cuh"):
    shutil.copy(os.path.join(csrc, f), d)
obj = os.path.join(d, "variant.o")
cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
       "-DFL_VARIANT=" + variant, "-DFL_DBLOCK_N=%d" % dblock, "-DFL_DMINB=%d" % minb, "-DFL_VBLOCK_N=%s" % os.environ.get("FL_VBLOCK", "128")] + extra + ["-c", os.path.join(d, "variant.cu"), "-o", obj]
r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
open(os.path.join(d, "build.log"), "w").write(r.stdout)
assert r.returncode == 0, r.stdout[-3000:]
import re
for m in re.finditer(r"k_node_derivsI\w+Li(\d)ELb(\d)ELb(\d)E.*?\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers", r.stdout):
    print("derivs<what %s, assemble %s, const-bank %s>: stack %s spill st %s ld %s regs %s" % m.groups())
std = os.path.join(csrc, "build")
stub = os.path.join(d, "stubs.cu")
with open(stub, "w") as fh:
    fh.write('struct fl_variant;\n' + "".join('extern "C" const fl_variant* fl_variant_%s() { return nullptr; }\n' % n for n in G.VARIANTS if n != variant))
subprocess.check_call(["nvcc", "-Xcompiler", "-fPIC", "-c", stub, "-o", os.path.join(d, "stubs.o")])
objs = [obj, os.path.join(d, "stubs.o"), os.path.join(std, "fl_capi.o"), os.path.join(std, "fl_ipm.o")]
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", os.path.join(d, "lib.so")] + objs + ["-lcudart"])

"""Attributes local-memory (spill) instructions of one kernel to source lines: python tools/sass_spills.py <object.o> <kernel-name-substring>"""
import collections, os, re, subprocess, sys, tempfile

obj, pat = sys.argv[1], sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout
cur, on = None, False
stl, ldl, ops = collections.Counter(), collections.Counter(), collections.Counter()
for ln in txt.splitlines():
    if ".section" in ln and ".text." in ln:
        on = pat in ln
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
    if m:
        op = m.group(2).split(".")[0]
        ops[op] += 1
        if op == "STL": stl[cur] += 1
        if op == "LDL": ldl[cur] += 1
print("instructions:", sum(ops.values()), dict(ops.most_common(14)))
print("STL by line:", sorted(stl.items(), key=lambda kv: -kv[1])[:10])
print("LDL by line:", sorted(ldl.items(), key=lambda kv: -kv[1])[:10])

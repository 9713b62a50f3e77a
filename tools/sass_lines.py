"""Static SASS census of one kernel in libpwb.so: how many instructions matching a regex (default: local
memory traffic, LDL/STL) come from each source line.
usage: sass_lines.py mangled_substring [opcode_regex] [top]"""
import os, re, subprocess, sys, tempfile
from collections import Counter
mangled = sys.argv[1]
pat = re.compile(sys.argv[2] if len(sys.argv) > 2 else r"\b(LDL|STL)\b")
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("PWB_SO", os.path.join(root, "pywindow_b200", "csrc", "libpwb.so"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
cub = [f for f in os.listdir(tmp) if f.startswith("pwb_kernels") and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
cur, infn, c, total, n = None, False, Counter(), 0, 0
for ln in dis.split("\n"):
    if ln.startswith(".text."):
        infn = mangled in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
    if m:
        n += 1
        if pat.search(m.group(2).replace(".", " ")):
            c[cur] += 1
            total += 1
print(f"{mangled}: {n} SASS instructions, {total} match {pat.pattern}")
src = {}
for (f, l), v in c.most_common(top):
    if f not in src:
        p = os.path.join(root, "pywindow_b200", "csrc", f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
    print("%4d  %s:%d  %s" % (v, f, l, text))

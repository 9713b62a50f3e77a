"""Per-source-line instruction counts / stall samples of one kernel: joins the SASS view of an .ncu-rep
(by instruction address) with nvdisasm line info of the cubin inside libpwb.so.
usage: ncu_lines.py report.ncu-rep kernel_regex mangled_substring [top] [sort column, e.g. stall_long_sb]"""
import csv, os, re, subprocess, sys, tempfile
from collections import Counter
rep, kern, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
sortcol = sys.argv[5] if len(sys.argv) > 5 else None
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.environ.get("PWB_SO", os.path.join(root, "pywindow_b200", "csrc", "libpwb.so"))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
unit = os.environ.get("PWB_CUBIN", "pwb_kernels")     # translation unit holding the kernel (pwb_kernels | pwb_periodic)
cub = [f for f in os.listdir(tmp) if f.startswith(unit + ".") and f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cub)], capture_output=True, text=True).stdout
lines, cur, infn, seq = dis.split("\n"), None, False, []
for ln in lines:
    if ln.startswith(".text."):
        infn = mangled in ln
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+\S", ln)
    if m:
        seq.append((int(m.group(1), 16), cur))
skip = os.environ.get("NCU_SKIP", "0")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern, "--launch-skip", skip, "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
end = next((i for i in range(2, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
data = [r for r in rows[2:end] if len(r) == len(hdr)]
ie, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
if sortcol:
    si = hdr.index(sortcol)
base = int(data[0][0], 16)
byaddr = dict(seq)
ci, cs = Counter(), Counter()
miss = 0
for r in data:
    off = int(r[0], 16) - base
    key = byaddr.get(off)
    if key is None:
        miss += 1
        key = ("?", 0)
    if r[ie].isdigit():
        ci[key] += int(r[ie]); cs[key] += int(r[si])
ti, ts = sum(ci.values()), sum(cs.values())
print(f"{kern}: {len(data)} SASS instr, {len(seq)} with line info, {miss} unmatched; dynamic {ti}")
src = {}
order = cs.most_common(top) if sortcol else ci.most_common(top)
for k, _ in order:
    v = ci[k]
    f, l = k
    if f not in src:
        p = os.path.join(root, "pywindow_b200", "csrc", f)
        src[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
    print("%5.1f%% inst %5.1f%% smp  %s:%d  %s" % (100 * v / ti, 100 * cs[k] / ts, f, l, text))

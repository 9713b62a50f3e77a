"""Hot-loop finder for an .ncu-rep: groups SASS instructions of one kernel by execution count."""
import csv, subprocess, sys
from collections import Counter
rep, kern = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
end = next((i for i in range(2, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
data = [r for r in rows[2:end] if len(r) == len(hdr)]
ie, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
cnt, smp, ex = Counter(), Counter(), {}
for r in data:
    if r[ie].isdigit():
        k = int(r[ie]); cnt[k] += 1; smp[k] += int(r[si]); ex.setdefault(k, []).append(r[1].strip())
tot = sum(k * v for k, v in cnt.items()); ts = sum(smp.values())
print(kern, "static", len(data), "dynamic warp-instr", tot)
for k, v in sorted(cnt.items(), key=lambda kv: -kv[0] * kv[1])[:8]:
    ops = Counter(x.split()[1].split(".")[0] if x.startswith("@") else x.split()[0].split(".")[0] for x in ex[k])
    print("  exec %10d x %4d instr = %5.1f%% inst %5.1f%% samples | %s" % (k, v, 100 * k * v / tot, 100 * smp[k] / ts,
          " ".join(f"{o}:{c}" for o, c in ops.most_common(8))))

#!/usr/bin/env python
"""Attribute the executed instructions / stall samples of one kernel in an .ncu-rep to CUDA source lines.

    python profiles/ncu_lines.py gpurun_out/prof.ncu-rep build/obj/amx_render.o k_tile [top]

ncu's CSV export of the source page carries metrics only for the SASS view; the line of every SASS instruction comes from
`nvdisasm -g` on the cubin inside the object file (compiled with -lineinfo).  Instructions are matched by order."""
import csv, io, os, re, subprocess, sys, tempfile
from collections import defaultdict

rep, obj, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
kname = rows[0][1]
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
sass = []
for r in rows[hi + 1:]:
    if not r or r[0] == "Kernel Name":
        break
    if len(r) == len(hdr):
        sass.append(r)
ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
# mangled name of the kernel: look it up through cuobjdump's symbol list
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = os.path.join(tmp, [f for f in os.listdir(tmp) if f.endswith(".cubin")][0])
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, text=True).stdout.splitlines()
dem = subprocess.run(["cu++filt"], input="\n".join(l for l in dis if l.startswith(".text.")), stdout=subprocess.PIPE, text=True).stdout.splitlines()
starts = [i for i, l in enumerate(dis) if l.startswith(".text.") and l.endswith(":")]
want = re.sub(r"\(bool\)|\(int\)|\s", "", kname.split("(amx::")[0])
sec = None
for i in starts:
    d = subprocess.run(["cu++filt", dis[i][6:-1]], stdout=subprocess.PIPE, text=True).stdout.strip()
    d2 = re.sub(r"\(bool\)|\(int\)|\s", "", d.split("(amx::")[0])
    if d2 == want:
        sec = i
        break
assert sec is not None, "kernel not found in " + obj
lines = []
cur = 0
for l in dis[sec + 1:]:
    if l.startswith(".text.") or l.startswith("//-------"):
        break
    m = re.search(r'//## File ".*?", line (\d+)', l)
    if m:
        # inlined code: keep the OUTERMOST?  nvdisasm prints the innermost location first, then 'inlined at'; take the first
        if "inlined at" not in l:
            cur = int(m.group(1))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
assert len(lines) == len(sass), (len(lines), len(sass))
inst, samp = defaultdict(float), defaultdict(float)
for ln, r in zip(lines, sass):
    inst[ln] += float(r[ci] or 0)
    samp[ln] += float(r[cs] or 0)
ti, ts = sum(inst.values()), sum(samp.values())
src = open(os.path.join(os.path.dirname(os.path.abspath(obj)), "..", "..", "atomorph_b200", "csrc", os.path.basename(obj).replace(".o", ".cu"))).read().splitlines()
print("# %s\n# %d SASS instructions, %.4g warp-instructions executed, %d stall samples" % (kname[:100], len(sass), ti, ts))
print("# line  %inst  %samples  source")
for ln in sorted(inst, key=lambda k: -(samp[k] if os.environ.get('BY_SAMPLES') else inst[k]))[:top]:
    print("%6d  %5.1f  %5.1f   %s" % (ln, 100 * inst[ln] / ti, 100 * samp[ln] / max(ts, 1), src[ln - 1].strip()[:120] if 0 < ln <= len(src) else "?"))

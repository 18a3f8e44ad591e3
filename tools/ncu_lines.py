"""Attribute ncu warp-stall samples (SASS page of an .ncu-rep captured with --import-source on, kernels built
with -lineinfo) to source lines, using nvdisasm's line annotations of the in-tree library.

    python tools/ncu_lines.py gpurun_out/r01c_step_full.ncu-rep k_step_full_v2ILi128ELb1 [top]
"""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, pattern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(root, "gym_narde_b200", "libnarde_b200.so")], cwd=tmp,
               stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
linemap, found = {}, False
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin") or "-" in f:
        continue
    out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    inside, cur = False, None
    for l in out.splitlines():
        if l.startswith("//--------------------- .text."):
            inside = pattern in l
            found |= inside
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", l)
        if m:
            linemap[int(m.group(1), 16)] = cur
assert found, "kernel not found in the library"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = [i for i, r in enumerate(rows) if len(r) > 3 and r[0] == "Address"][0]
hdr = rows[hi]
sa, ie, te = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
samp = []
for r in rows[hi + 1:]:          # the first launch of the report only (a report may hold several)
    if len(r) <= te or not r[0].startswith("0x"):
        break
    if r[sa].isdigit():
        samp.append((int(r[0], 16), int(r[sa]), int(r[ie]), int(r[te])))
base = samp[0][0]
agg = collections.defaultdict(lambda: [0, 0, 0])
tot, toti = sum(s[1] for s in samp), sum(s[2] for s in samp)
for a, s, i, t in samp:
    k = linemap.get(a - base)
    agg[k][0] += s; agg[k][1] += i; agg[k][2] += t
print("samples", tot, "warp instructions", toti)
for k, (s, i, t) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%6d %5.1f%%  inst %5.1f%%  lanes %4.1f  %s" % (s, 100 * s / tot, 100 * i / max(toti, 1), t / max(i, 1), "%s:%d" % k if k else "?"))

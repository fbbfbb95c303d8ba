#!/usr/bin/env python
"""Join an ncu SASS-level source page (ncu -i rep --page source --csv) with nvdisasm -gi line info of the same cubin and
aggregate executed instructions / stall samples by the OUTERMOST source line in the kernel's file.
usage: ncu_by_line.py src.csv disasm_gi.txt mangled_kernel_name kernel_file_basename [top]"""
import csv, re, sys, collections
src, dis, kname, kfile = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 60
# ---- address -> (outer line in kfile, innermost file:line)
addr2loc = {}
infn = False
group = []
fresh = False
for ln in open(dis):
    if ln.startswith("//---") and ".text." in ln:
        infn = (".text." + kname + " ") in ln
        group = []
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        if not fresh:
            group = []
            fresh = True
        group.append((m.group(1), int(m.group(2))))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
    if m:
        fresh = False
        a = int(m.group(1), 16)
        outer = None
        for f, l in group:
            if f.endswith(kfile):
                outer = l
        inner = group[0] if group else ("?", 0)
        addr2loc[a] = (outer, inner, m.group(2))
rows = list(csv.reader(open(src)))
hdr = None
for i, r in enumerate(rows):
    if r and r[0] == "Address":
        hdr = r; start = i + 1; break
ix = {n: hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_no_inst", "stall_long_sb", "stall_wait", "stall_short_sb", "stall_math", "stall_lg", "L2 Theoretical Sectors Local")}
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0, 0, 0])
base = None
tot = [0, 0, 0]
op = collections.defaultdict(lambda: [0, 0])
for r in rows[start:]:
    if len(r) < len(hdr):
        continue
    a = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]])
    if base is None:
        base = a
    off = a - base
    f = lambda n: float(r[ix[n]] or 0)
    outer, inner, ins = addr2loc.get(off, (None, ("?", 0), "?"))
    key = outer
    v = agg[key]
    v[0] += f("Instructions Executed"); v[1] += f("Thread Instructions Executed"); v[2] += f("# Samples")
    v[3] += f("stall_no_inst"); v[4] += f("stall_long_sb"); v[5] += f("L2 Theoretical Sectors Local"); v[6] += 1
    tot[0] += f("Instructions Executed"); tot[1] += f("Thread Instructions Executed"); tot[2] += f("# Samples")
    o = ins.split()[0] if not ins.startswith("@") else ins.split()[1]
    o = o.split(".")[0]
    op[o][0] += f("Instructions Executed"); op[o][1] += f("# Samples")
lines = open("/root/repo/ptmcmc_b200/csrc/" + kfile).read().split("\n")
print("total warp-inst %.4g thread-inst %.4g samples %d  avg lanes %.1f" % (tot[0], tot[1], tot[2], tot[1] / tot[0]))
print("%6s %7s %7s %6s %6s %6s %5s  %s" % ("line", "inst%", "samp%", "lanes", "noinst", "longsb", "nSASS", "source"))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    s = lines[k - 1].strip()[:110] if k else "(no line)"
    print("%6s %7.2f %7.2f %6.1f %6.1f %6.1f %5d  %s" % (k, 100 * v[0] / tot[0], 100 * v[2] / tot[2], v[1] / max(v[0], 1), 100 * v[3] / max(v[2], 1), 100 * v[4] / max(v[2], 1), v[6], s))
print("\nby opcode (inst%, samples%):")
for k, v in sorted(op.items(), key=lambda kv: -kv[1][0])[:25]:
    print("  %-10s %6.2f %6.2f" % (k, 100 * v[0] / tot[0], 100 * v[1] / tot[2]))

"""Development helper: join an ncu SASS source page (csv) with nvdisasm -g line info and
aggregate stall samples per source line.  usage: ncu_lines.py sass.csv disasm.txt kernel_substr"""
import csv, re, sys, collections
sass_csv, dis, kname = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: address -> (file, line) inside the section of the kernel
addr2line = {}
insec = False; cur = ("?", 0)
for ln in open(dis):
    if ln.startswith("//-----") and ".text." in ln:
        insec = kname in ln
        continue
    if not insec: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]+)\*/\s+(\S.*);", ln)
    if m:
        addr2line[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
base = None
agg = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, 0])
for r in rows[2:]:
    a = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]], 16)
    if base is None: base = a
    key = addr2line.get(a - base, ("?", 0))
    g = agg[key]
    g[0] += f(r, "# Samples"); g[1] += f(r, "Instructions Executed"); g[2] += f(r, "Thread Instructions Executed")
    g[3] += f(r, "stall_no_inst"); g[4] += 1
tot = sum(g[0] for g in agg.values()); ti = sum(g[1] for g in agg.values())
print("total samples %.0f, warp instr %.3g, static instr %d" % (tot, ti, sum(g[4] for g in agg.values())))
byfile = collections.defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, 0])
for (fl, ln), g in agg.items():
    for k in range(5): byfile[fl][k] += g[k]
for fl, g in sorted(byfile.items(), key=lambda kv: -kv[1][0]):
    print("%-24s samples %5.1f%%  instr %5.1f%%  thr/instr %4.1f  no_inst %5.1f%%  static %5d" % (fl, 100 * g[0] / tot, 100 * g[1] / ti, g[2] / max(g[1], 1), 100 * g[3] / max(g[0], 1), g[4]))
print()
for (fl, ln), g in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%-22s:%4d samples %5.2f%%  instr %5.2f%%  thr/instr %4.1f  no_inst %4.0f%%  static %4d" % (fl, ln, 100 * g[0] / tot, 100 * g[1] / ti, g[2] / max(g[1], 1), 100 * g[3] / max(g[0], 1), g[4]))

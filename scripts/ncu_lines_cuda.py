"""Development helper: samples per CUDA source line of an ncu report (needs -lineinfo and --import-source on).
usage: ncu_lines_cuda.py report.ncu-rep [top]"""
import csv, subprocess, sys, io, collections
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(io.StringIO(raw)))
cur, hdr = None, None
lines = []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0] not in ("", "Function Name") and cur:
        ix = {h: k for k, h in enumerate(hdr)}
        try:
            smp = int(float(r[ix["# Samples"]] or 0))
        except (ValueError, IndexError):
            continue
        st = {h[6:]: int(float(r[k] or 0)) for k, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h and k < len(r)}
        ie = int(float(r[ix["Instructions Executed"]] or 0)); te = int(float(r[ix["Thread Instructions Executed"]] or 0))
        lines.append((smp, cur, r[0], r[1].strip()[:70], st, ie, te))
tot = sum(l[0] for l in lines)
byfile = collections.Counter()
for l in lines:
    byfile[l[1]] += l[0]
print("samples", tot, dict(byfile))
for smp, f, ln, src, st, ie, te in sorted(lines, reverse=True)[:top]:
    t3 = sorted(st.items(), key=lambda x: -x[1])[:3]
    print("%5.2f%% %-20s %5s lanes %4.1f  %-70s %s" % (100 * smp / tot, f, ln, te / max(ie, 1), src, " ".join("%s:%d%%" % (k, 100 * v / max(smp, 1)) for k, v in t3)))
# attribution of samples to regions of the Gaussian kernels (file, first line, last line, name)
REG = [("pfx_lane_lm.cuh", 34, 116, "qrfac_small"), ("pfx_lane_lm.cuh", 167, 230, "exp / rcp / rsqrt"),
       ("pfx_lane_lm.cuh", 231, 438, "SmallLM lmpar"), ("pfx_lane_lm.cuh", 688, 761, "jac_end"),
       ("pfx_lane_lm.cuh", 762, 817, "trial_begin"), ("pfx_lane_lm.cuh", 851, 905, "trial_end"),
       ("pfx_lane_lm.cuh", 489, 520, "lm init/jac_begin")]
if len(sys.argv) > 3:
    agg = collections.Counter(); lanes = collections.defaultdict(lambda: [0, 0])
    for smp, f, ln, src, st, ie, te in lines:
        name = f
        for rf, a, b, nm in REG:
            if f == rf and a <= int(ln) <= b:
                name = nm
        if f in ("pfx_lane_gauss.cuh", "pfx_k_gauss.cu"):
            name = f + ":" + str(int(ln) // 50 * 50)
        agg[name] += smp; lanes[name][0] += ie; lanes[name][1] += te
    for k, v in agg.most_common(40):
        print("%5.1f%%  lanes %4.1f  %s" % (100 * v / tot, lanes[k][1] / max(lanes[k][0], 1), k))

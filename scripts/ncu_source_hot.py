"""Aggregate the SASS source page of an ncu report: samples by opcode, stall reasons, hot windows.
usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python scripts/ncu_source_hot.py src.csv [window]"""
import csv, re, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
W = int(sys.argv[2]) if len(sys.argv) > 2 else 200
kernels = []
i = 0
while i < len(rows):
    if rows[i] and rows[i][0] == 'Kernel Name':
        name = rows[i][1]; hdr = rows[i + 1]; i += 2; data = []
        while i < len(rows) and not (rows[i] and rows[i][0] == 'Kernel Name'):
            if len(rows[i]) >= len(hdr): data.append(rows[i])
            i += 1
        kernels.append((name, hdr, data))
    else:
        i += 1
for name, hdr, data in kernels:
    ix = {h: k for k, h in enumerate(hdr)}
    S = lambda r, k: int(float(r[ix[k]] or 0))
    tot = sum(S(r, '# Samples') for r in data)
    print('==', name[:60], 'samples', tot, 'sass instrs', len(data))
    op = collections.Counter(); opi = collections.Counter(); opt = collections.Counter()
    for r in data:
        m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']])
        o = m.group(2).split('.')[0] if m else '?'
        op[o] += S(r, '# Samples'); opi[o] += S(r, 'Instructions Executed'); opt[o] += S(r, 'Thread Instructions Executed')
    ti = sum(opi.values())
    print('warp instr', ti, 'lanes/instr %.2f' % (sum(opt.values()) / ti))
    for o, c in op.most_common(22):
        print('  %-10s samples %5.1f%%  instr %5.1f%%  lanes %.1f' % (o, 100 * c / tot, 100 * opi[o] / ti, opt[o] / max(1, opi[o])))
    st = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    tots = {h: sum(S(r, h) for r in data) for h in st}
    print('  stalls:', ', '.join('%s %.1f%%' % (k[6:], 100 * v / tot) for k, v in sorted(tots.items(), key=lambda x: -x[1])[:8]))
    for a in range(0, len(data), W):
        blk = data[a:a + W]
        s = sum(S(r, '# Samples') for r in blk); ie = sum(S(r, 'Instructions Executed') for r in blk); te = sum(S(r, 'Thread Instructions Executed') for r in blk)
        if s > 0.015 * tot:
            tops = collections.Counter()
            for r in blk:
                for h in st: tops[h[6:]] += S(r, h)
            print('  sass %5d-%5d samples %5.1f%% exec %5.1f%% lanes %4.1f rel-cost %.2f  %s' % (
                a, a + W, 100 * s / tot, 100 * ie / ti, te / max(ie, 1), (s / tot) / (ie / ti) if ie else 0,
                ' '.join('%s:%d%%' % (k, 100 * v / s) for k, v in tops.most_common(3))))

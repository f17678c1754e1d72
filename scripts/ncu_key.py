"""Development helper: key metrics and the stall breakdown of every kernel in an .ncu-rep (ncu --page raw --csv).
usage: ncu_key.py report.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
pick = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'sm__warps_active.avg.per_cycle_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__inst_executed.sum', 'smsp__average_warp_latency_per_inst_issued.ratio',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__sass_inst_executed_op_local_ld.sum',
        'smsp__sass_inst_executed_op_local_st.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__icc_requests_lookup_hit.sum', 'sm__icc_requests_lookup_miss.sum']
for vals in rows[2:]:
    print("==", vals[hdr.index("Kernel Name")][:70])
    for k in pick:
        if k in hdr:
            i = hdr.index(k)
            print("  %-70s %s %s" % (k, vals[i], units[i]))
    st = []
    for i, h in enumerate(hdr):
        if 'average_warp' in h and 'issue_stalled' in h and h.endswith('_per_issue_active.ratio') and 'not_issued' not in h:
            try:
                st.append((float(vals[i].replace(',', '')), h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', '')))
            except ValueError:
                pass
    print("  stalls (cycles per issued instruction):", ", ".join("%s %.2f" % (h, v) for v, h in sorted(st, reverse=True)[:9]))

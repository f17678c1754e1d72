"""Turns the ncu metric list of the two Gaussian kernels (scripts/calls/call32.sh: FP64 thread instructions,
DRAM bytes, duration; 100 000 HTRU candidates) into the two small JSON files bench.py reads for its
roofline entries: profiles/r02_fp64_flops.json and profiles/r02_gauss_summary.json.
usage: ncu_gauss_json.py metrics.csv n_candidates"""
import csv, json, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0] != "ID"]
n = int(sys.argv[2])
per = {}
for r in rows:
    k = r[4].split("(")[0].split("::")[-1]
    per.setdefault(k, {})[r[12]] = float(r[14].replace(",", ""))
out = {"source": "profiles/r02_fp64_ops.csv (ncu --metrics smsp__sass_thread_inst_executed_op_{dfma,dadd,dmul}_pred_on.sum, "
                 "dram__bytes_{read,write}.sum, %d HTRU candidates, scripts/calls/call32.sh)" % n, "candidates": n}
tot_flop, tot_dram = 0.0, 0.0
for k, m in per.items():
    flop = 2 * m.get("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", 0) + \
        m.get("smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", 0) + m.get("smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", 0)
    dram = m.get("dram__bytes_read.sum", 0) + m.get("dram__bytes_write.sum", 0)
    ms = m.get("gpu__time_duration.sum", 0) / 1e6
    out[k] = {"fp64_flop": flop, "ms_under_ncu": ms, "tflops_under_ncu": flop / (ms / 1e3) / 1e12 if ms else None,
              "dram_bytes": dram}
    tot_flop += flop
    tot_dram += dram
out["gauss_pair_fp64_flop_per_candidate"] = tot_flop / n
json.dump(out, open("profiles/r02_fp64_flops.json", "w"), indent=1)
json.dump({"source": "profiles/r02_fp64_ops.csv (ncu dram__bytes_read.sum + dram__bytes_write.sum of k_gauss_rounds_s + k_gauss_final_s, %d candidates)" % n,
           "dram_bytes_per_candidate": tot_dram / n}, open("profiles/r02_gauss_summary.json", "w"), indent=1)
print(json.dumps(out, indent=1))

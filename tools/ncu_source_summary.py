"""Summarise an ncu report exported with
     ncu -i X.ncu-rep --page raw --csv > raw.csv ; ncu -i X.ncu-rep --page source --csv --print-source sass > sass.csv
into the handful of numbers DESIGN.md / profiles/ quote."""
import csv, collections, re, sys
raw, sass = sys.argv[1], sys.argv[2]
rows = list(csv.reader(open(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__average_warp_latency_per_inst_issued.ratio", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.avg", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum"]
for i, h in enumerate(hdr):
    if h in want or ("stalled" in h and "per_issue_active" in h and "not_issued" not in h):
        try:
            v = float(vals[i].replace(',', ''))
            if "stalled" in h and v < 0.05:
                continue
        except ValueError:
            pass
        print(h, vals[i], units[i])
rows = list(csv.reader(open(sass)))
print("kernel:", rows[0][1][:90])
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
tot = 0; by = collections.Counter(); samp = collections.Counter(); ts = 0; thr = 0
sh = []
for r in data:
    if len(r) < len(hdr):
        continue
    src = r[ix["Source"]]; ex = int(r[ix["Instructions Executed"]] or 0); s = int(r[ix["# Samples"]] or 0)
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", src)
    op = m.group(2).split('.')[0] if m else '?'
    by[op] += ex; samp[op] += s; tot += ex; ts += s; thr += int(r[ix["Thread Instructions Executed"]] or 0)
    if r[ix["L1 Wavefronts Shared"]] not in ("", "0"):
        sh.append((int(r[ix["L1 Wavefronts Shared"]]), int(r[ix["L1 Wavefronts Shared Ideal"]] or 0), src.strip()))
print("warp instructions", tot, "thread instructions", thr, "samples", ts)
print("executed by opcode:", [(k, round(v / tot, 3)) for k, v in by.most_common(24)])
print("stall samples by opcode:", [(k, round(v / ts, 3)) for k, v in samp.most_common(12)])
fp64 = sum(by[k] for k in ("DFMA", "DMUL", "DADD", "DSETP"))
print("FP64 share of issued instructions: %.3f" % (fp64 / tot))
print("shared wavefronts", sum(a for a, b, c in sh), "ideal", sum(b for a, b, c in sh))
for a, b, c in sorted(sh, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 12]:
    print("  ", a, b, round(a / max(b, 1), 2), c)

"""Summarise an .ncu-rep by CUDA source line: python tools/ncu_lines.py rep [top]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'launch__grid_size', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'sm__inst_executed_pipe_lsu.sum', 'smsp__inst_executed_pipe_fp64.sum', 'sm__sass_inst_executed_op_shared.sum']
for w in want:
    for i, h in enumerate(hdr):
        if h == w:
            print(w, rows[1][i], [r[i] for r in rows[2:]])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
iSamp = hdr.index("# Samples"); iInst = hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
lines = []; tot_s = tot_i = 0; agg = {}
for r in rows[hi + 1:]:
    if r and r[0] != '':
        try:
            s = int(r[iSamp]); n = int(r[iInst])
        except Exception:
            continue
        st = {hdr[i]: int(r[i] or 0) for i in stall_cols}
        for k, v in st.items():
            agg[k] = agg.get(k, 0) + v
        lines.append((s, n, int(r[0]), r[1].strip()[:100], st)); tot_s += s; tot_i += n
print('total samples', tot_s, 'total inst', tot_i)
print([(k, round(100 * v / tot_s, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:9]])
for s, n, l, t, st in sorted(lines, key=lambda x: -x[0])[:top]:
    tp = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print('%5.1f%% smp %5.1f%% inst  L%-4d %s   %s' % (100 * s / tot_s, 100 * n / tot_i, l, t, [(k[6:], v) for k, v in tp]))

"""Summarise an .ncu-rep: key metrics, stall breakdown, hottest SASS lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__grid_size', 'launch__block_size', 'sm__cycles_elapsed.avg',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_warps', 'smsp__warps_eligible.avg.per_cycle_active', 'launch__waves_per_multiprocessor',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.sum', 'lts__t_bytes.sum',
        'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum', 'launch__shared_mem_per_block_dynamic']
for r in rows[2:]:
    for k in keys:
        if k in hdr:
            print(f"{k} = {r[hdr.index(k)][:90]} {units[hdr.index(k)]}")
    d = {}
    for i, k in enumerate(hdr):
        if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k:
            try:
                d[k.replace('smsp__pcsamp_warps_issue_stalled_', '')] = float(r[i].replace(',', ''))
            except Exception:
                pass
    tot = sum(d.values()) or 1
    print("stalls:", ", ".join(f"{k} {100 * v / tot:.1f}%" for k, v in sorted(d.items(), key=lambda x: -x[1])[:8]))
    print('---')
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
body = [r for r in rows[2:] if len(r) == len(h)]
si, sc, ie = h.index('# Samples'), h.index('Source'), h.index('Instructions Executed')
def num(x):
    try: return int(x)
    except Exception: return 0
idx = sorted(range(len(body)), key=lambda i: -num(body[i][si]))[:topn]
print("total samples", sum(num(r[si]) for r in body), "sass lines", len(body))
for i in sorted(idx):
    r = body[i]
    st = {k: num(r[h.index(k)]) for k in h if k.startswith('stall_') and 'Not Issued' not in k}
    top = sorted(st.items(), key=lambda x: -x[1])[:2]
    print(i, r[si], r[ie], r[sc].strip()[:72], '|', ' '.join(f"{k[6:]}={v}" for k, v in top if v))

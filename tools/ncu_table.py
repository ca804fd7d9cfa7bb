"""One line per profiled launch of an .ncu-rep: time, DRAM bytes/throughput, pipes, occupancy, top stalls."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
def col(r, k, d=0.0):
    if k not in hdr: return d
    try: return float(r[hdr.index(k)].replace(',', ''))
    except Exception: return d
print("idx,kernel,grid,block,regs,us,dram_rd_MB,dram_wr_MB,dram_pct,l2_pct,l1_pct,issue_pct,fma_pct,alu_pct,xu_pct,lsu_pct,tensor_pct,warps_active_pct,top_stalls")
for i, r in enumerate(rows[2:]):
    name = r[hdr.index('Kernel Name')][:44].replace(',', ';')
    st = {}
    for j, k in enumerate(hdr):
        if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k:
            try: st[k.replace('smsp__pcsamp_warps_issue_stalled_', '')] = float(r[j].replace(',', ''))
            except Exception: pass
    tot = sum(st.values()) or 1
    top = " ".join(f"{k}:{100*v/tot:.0f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:4])
    u = rows[1]
    def mb(k):
        v = col(r, k); unit = u[hdr.index(k)] if k in hdr else ''
        return v * {'byte': 1e-6, 'Kbyte': 1e-3, 'Mbyte': 1, 'Gbyte': 1e3}.get(unit, 1)
    def us(k):
        v = col(r, k); unit = u[hdr.index(k)]
        return v * {'ns': 1e-3, 'us': 1, 'ms': 1e3, 's': 1e6}.get(unit, 1)
    print(f"{i},{name},{r[hdr.index('launch__grid_size')]},{r[hdr.index('launch__block_size')]},{r[hdr.index('launch__registers_per_thread')]},"
          f"{us('gpu__time_duration.sum'):.1f},{mb('dram__bytes_read.sum'):.1f},{mb('dram__bytes_write.sum'):.1f},"
          f"{col(r,'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f},{col(r,'lts__throughput.avg.pct_of_peak_sustained_elapsed'):.1f},"
          f"{col(r,'l1tex__throughput.avg.pct_of_peak_sustained_elapsed'):.1f},{col(r,'smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f},"
          f"{col(r,'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):.1f},{col(r,'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active'):.1f},"
          f"{col(r,'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):.1f},{col(r,'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):.1f},"
          f"{col(r,'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f},{col(r,'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f},{top}")

"""Summarise an .ncu-rep here (no GPU): headline raw metrics + stall reasons + hottest SASS lines."""
import csv, subprocess, sys, io
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h = r[0]; v = r[2] if len(r) > 2 else r[1]
for k in ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
          'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sector_hit_rate.pct',
          'l1tex__m_xbar2l1tex_read_bytes.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
          'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_elapsed.max']:
    if k in h:
        print("%-70s %s %s" % (k, v[h.index(k)], r[1][h.index(k)] if len(r) > 2 else ""))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; idx = {k: i for i, k in enumerate(hdr)}
data = rows[2:]
tot = sum(int(x[idx['# Samples']]) for x in data)
print("total samples", tot)
reasons = [k for k in hdr if k.startswith('stall_') and 'Not Issued' not in k]
agg = {k: sum(int(x[idx[k]] or 0) for x in data) for k in reasons}
for k, val in sorted(agg.items(), key=lambda x: -x[1])[:8]:
    print("  %-28s %8d %5.1f%%" % (k, val, 100 * val / tot))
for x in sorted(data, key=lambda x: -int(x[idx['# Samples']]))[:ntop]:
    st = {k: int(x[idx[k]] or 0) for k in reasons}
    dom = sorted(st.items(), key=lambda y: -y[1])[:2]
    print(x[idx['# Samples']].rjust(7), x[idx['Instructions Executed']].rjust(9), x[idx['Source']].strip()[:64].ljust(64), dom)

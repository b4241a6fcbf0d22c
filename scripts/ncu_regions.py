"""Folds an `ncu --page source --csv` export into blocks of SASS instructions: executed instructions per
warp-tile, share of samples and the top stall reasons per block (used to read the evaluation kernel's profiles)."""
import csv, sys
raw, src = sys.argv[1], sys.argv[2]
tiles = float(sys.argv[3]) if len(sys.argv) > 3 else 9.77e7
step = int(sys.argv[4]) if len(sys.argv) > 4 else 25
rows = list(csv.reader(open(raw)))
d = dict(zip(rows[0], rows[2]))
for k in ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
          'smsp__warps_eligible.avg.per_cycle_active', 'smsp__warps_active.avg.per_cycle_active', 'launch__registers_per_thread',
          'dram__bytes_read.sum', 'dram__bytes_write.sum']:
    print(k, d.get(k))
for k in rows[0]:
    if 'smsp__average_warps_issue_stalled' in k and k.endswith('ratio') and float(d[k] or 0) > 0.2:
        print(' ', k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), d[k])
rows = list(csv.reader(open(src)))
hdr = rows[1]
isrc = hdr.index('Source'); iex = hdr.index('Instructions Executed'); ismp = hdr.index('# Samples')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data = [r for r in rows[2:] if len(r) > iex]
tot = sum(int(r[iex] or 0) for r in data); tots = sum(int(r[ismp] or 0) for r in data)
print('total instr', tot, 'samples', tots)
for i in range(0, len(data), step):
    blk = data[i:i + step]
    e = sum(int(b[iex] or 0) for b in blk); sm = sum(int(b[ismp] or 0) for b in blk)
    if e / tot > 0.004 or sm / tots > 0.006:
        st = {hdr[c]: sum(int(b[c] or 0) for b in blk) for c in stall_cols}
        top = sorted(st.items(), key=lambda x: -x[1])[:3]
        print(f"{i:5d} {e / tiles:7.2f}/wt {100 * e / tot:5.1f}%i {100 * sm / tots:5.1f}%s  {blk[0][isrc][:42]:42s} "
              f"{[(k[6:], round(100 * v / max(sm, 1))) for k, v in top]}")

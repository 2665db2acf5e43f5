"""Summarise an .ncu-rep here (no GPU): key raw metrics + hot SASS regions.  usage: ncu_summary.py rep [kernel-idx]"""
import csv, subprocess, sys, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keep = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__block_size', 'launch__shared_mem_per_block_allocated', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed']
for vals in rows[2:]:
    for h, u, v in zip(hdr, units, vals):
        if h in keep or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') and float(v or 0) > 0.3):
            print(f'{h},{u},{v}')
    print('----')
if '--src' in sys.argv:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]; data = rows[2:]
    si = hdr.index('Warp Stall Sampling (All Samples)'); ii = hdr.index('Instructions Executed'); ti = hdr.index('Avg. Threads Executed')
    tot_s = sum(int(r[si]) for r in data); tot_i = sum(int(r[ii]) for r in data)
    print('samples', tot_s, 'inst', tot_i, 'sass lines', len(data))
    for n, r in enumerate(data):
        s_, i_ = int(r[si]), int(r[ii])
        if s_ > 0.01 * tot_s or i_ > 0.015 * tot_i:
            print(f"{n:4d} {r[1][:66]:66s} samp {100*s_/tot_s:5.1f}% inst {100*i_/tot_i:5.1f}% thr {r[ti]}")

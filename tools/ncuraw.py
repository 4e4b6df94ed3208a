"""Print the handful of raw ncu metrics we read from a .ncu-rep (development aid). usage: python tools/ncuraw.py file.ncu-rep"""
import csv, sys, subprocess
f = sys.argv[1]
out = subprocess.run(['ncu', '-i', f, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__maximum_warps_per_active_cycle_pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
idx = {h: i for i, h in enumerate(hdr)}
stall = [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' not in h and h.endswith('.ratio')]
for r in rows[2:]:
    print('---')
    for k in keys:
        if k in idx:
            print(' ', k, '=', r[idx[k]], rows[1][idx[k]])
    st = sorted(((float(r[idx[h]] or 0), h) for h in stall), reverse=True)[:6]
    for v, h in st:
        print('   stall', h.replace('smsp__average_warp_latency_issue_stalled_', '').replace('smsp__average_warps_issue_stalled_', ''), round(v, 2))

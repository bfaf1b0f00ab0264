"""Print the metrics we track from an `ncu --page raw --csv` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'launch__waves_per_multiprocessor', 'smsp__issue_active.avg.pct', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio']
idx = {h: i for i, h in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith('smsp__average_warp') and 'issue_stalled' in h and h.endswith('_per_issue_active.ratio')] or \
         [h for h in hdr if 'issue_stalled' in h and h.endswith('.pct')]
for r in rows[2:]:
    print('----')
    for w in want:
        if w in idx:
            print(f'{w:75s} {r[idx[w]]}')
    st = []
    for h in stalls:
        try:
            st.append((float(r[idx[h]].replace(',', '')), h))
        except ValueError:
            pass
    for v, h in sorted(st, reverse=True)[:7]:
        print(f'   stall {h:90s} {v:.2f}')

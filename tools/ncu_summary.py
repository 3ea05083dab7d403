"""Compact summary of an .ncu-rep (run where ncu is installed; no GPU needed): key metrics + top stall sites."""
import csv, subprocess, sys, io, re
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'sm__cycles_elapsed.avg',
        'sm__cycles_elapsed.avg.per_second', 'smsp__inst_executed.sum', 'sm__warps_active.avg.per_cycle_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'lts__t_bytes.sum', 'l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_barrier', 'smsp__pcsamp_warps_issue_stalled_selected', 'smsp__pcsamp_warps_issue_stalled_wait',
        'smsp__pcsamp_warps_issue_stalled_no_instructions', 'smsp__pcsamp_warps_issue_stalled_short_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_not_selected',
        'smsp__pcsamp_warps_issue_stalled_mio_throttle', 'smsp__pcsamp_warps_issue_stalled_lg_throttle',
        'smsp__pcsamp_warps_issue_stalled_dispatch_stall', 'smsp__pcsamp_warps_issue_stalled_branch_resolving',
        'launch__grid_size', 'launch__block_size']
for r in rows[2:]:
    print('== kernel', r[hdr.index('Kernel Name')][:60] if 'Kernel Name' in hdr else '')
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f'  {k:85s} {r[i]:>16s} {units[i]}')
src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; data = rows[2:]
si = h.index('Warp Stall Sampling (All Samples)')
tot = sum(int(r[si] or 0) for r in data)
print('total stall samples', tot, ' sass instrs', len(data))
# annotate each instruction with the nearest preceding "marker" (LDTM/STTM/UTC*/SYNCS/BAR) to get a feel for the phase
top = sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[:ntop]
for i in top:
    print(f'{int(data[i][si]):6d} {100*int(data[i][si])/max(tot,1):5.1f}%  #{i:5d} {data[i][1].strip()[:80]}')

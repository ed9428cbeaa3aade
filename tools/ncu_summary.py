# Summarises an .ncu-rep (read on the CPU box): per-launch headline metrics, stall reasons and the hottest SASS regions.
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__cycles_active.avg', 'sm__cycles_elapsed.avg',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed', 'idc__requests.sum', 'idc__requests_lookup_miss.sum',
        'sm__inst_executed_pipe_uniform.sum', 'smsp__inst_executed_op_ldc.sum']
for r in rows[2:]:
    for w in want:
        if w in hdr:
            print(f"{w:80s} {r[hdr.index(w)]} {rows[1][hdr.index(w)]}")
    for i, h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                if float(r[i]) > 0.15:
                    print('   stall', h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), r[i])
            except ValueError:
                pass
    print('---')
if len(sys.argv) > 2:
    which = int(sys.argv[2])
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    # the source page holds one table per launch, separated by a "Kernel Name" line
    blocks = src.split('"Kernel Name"')
    blk = blocks[which + 1]
    rows = list(csv.reader(io.StringIO('"Kernel Name"' + blk)))
    hdr = rows[1]
    ia, isrc, ie = hdr.index('Address'), hdr.index('Source'), hdr.index('Instructions Executed')
    ist = hdr.index('Warp Stall Sampling (All Samples)')
    data = [(r[ia], r[isrc], int(r[ie] or 0), int(r[ist] or 0)) for r in rows[2:] if len(r) > ist and r[ie].isdigit()]
    tot = sum(d[2] for d in data)
    tots = sum(d[3] for d in data)
    print('total warp instr', tot, 'sass lines', len(data), 'samples', tots)
    ops = collections.Counter()
    for a, s, e, st in data:
        t = s.split()
        ops[t[1] if t[0].startswith('@') else t[0]] += e
    for k, v in ops.most_common(22):
        print(f"  {k:34s} {v:12d} {100 * v / tot:5.1f}%")
    print('top stall-sample instructions:')
    for i in sorted(range(len(data)), key=lambda i: -data[i][3])[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
        a, s, e, st = data[i]
        print(f"  #{i:5d} {100 * st / max(tots, 1):5.1f}%  exec {e:10d}  {s}")

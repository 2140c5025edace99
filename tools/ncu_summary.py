"""Summarise an .ncu-rep: key raw metrics per kernel, and per-source-line sample / instruction shares."""
import csv, subprocess, sys, collections, io

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'smsp__inst_executed.sum',
        'sm__issue_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct', 'sm__cycles_active.avg', 'launch__waves_per_multiprocessor']
want += [h for h in hdr if h.startswith('smsp__pcsamp_warps_issue_stalled') and 'not_issued' not in h]
for k, r in enumerate(rows[2:]):
    if which is not None and k != which:
        continue
    print(f"=== launch {k}")
    for i, h in enumerate(hdr):
        if h in want:
            print(f"  {h:85s} {units[i]:10s} {r[i][:90]}")
if which is not None:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    # the source page lists kernels one after another; take the block of launch `which`
    blocks = src.split('"File Path"')
    blk = '"File Path"' + blocks[1 + which] if len(blocks) > 1 + which else src
    rows = list(csv.reader(io.StringIO(blk)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Line No')
    hdr = rows[hi]
    iS, iI = hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
    agg = collections.OrderedDict(); ts = ti = 0
    for r in rows[hi + 1:]:
        if r and r[0] not in ('', 'Line No'):
            try: s, n = int(r[iS]), int(r[iI])
            except Exception: continue
            a = agg.setdefault(int(r[0]), [r[1].strip()[:100], 0, 0]); a[1] += s; a[2] += n; ts += s; ti += n
    print('total samples', ts, 'total warp-inst', ti)
    for ln, (s_, s, n) in sorted(agg.items()):
        if s / max(ts, 1) > 0.006 or n / max(ti, 1) > 0.006:
            print(f"{ln:4d} samp={s/ts:6.1%} inst={n/ti:6.1%}  {s_}")

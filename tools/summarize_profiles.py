"""Turn gpurun_out/ captures (ncu launch list CSV, --set full .ncu-rep files, phase stamps) into the
text summaries committed under profiles/.  Usage: python tools/summarize_profiles.py r1"""
import collections, csv, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else 'r1'
src, dst = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
os.makedirs(dst, exist_ok=True)

WANT = ('gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_tensor',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__inst_executed.sum', 'sm__cycles_elapsed.max', 'sm__inst_executed_pipe_xu.sum', 'smsp__cycles_active.avg',
        'smsp__average_warp_latency_issue_stalled', 'smsp__average_warps_issue_stalled', 'sm__cycles_active.avg',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct')


def launches(path, out):
    lines = [l for l in open(path) if not l.startswith('==')]
    rows = list(csv.DictReader(lines))
    agg, tot = collections.defaultdict(lambda: [0, 0.0]), 0.0
    for r in rows:
        v = float(r['Metric Value'].replace(',', '')) / 1000.0
        agg[r['Kernel Name']][0] += 1
        agg[r['Kernel Name']][1] += v
        tot += v
    with open(out, 'w') as f:
        f.write(f'# {os.path.basename(path)}: {len(rows)} launches, {tot:.0f} us of device time (ncu: cold cache, serialised)\n')
        f.write('#   total_us  share  count   avg_us  kernel\n')
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f'{t:10.1f} {100 * t / tot:6.2f}% {n:6d} {t / n:9.2f}  {k[:160]}\n')


def full(path, out):
    raw = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(out, 'w') as f:
        for vals in rows[2:]:
            name = vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '?'
            f.write(f'## {name}\n')
            for i, h in enumerate(hdr):
                if any(h.startswith(w) for w in WANT):
                    f.write(f'{h:78s} {units[i]:16s} {vals[i]}\n')
        det = subprocess.run(['ncu', '-i', path, '--page', 'details'], capture_output=True, text=True).stdout
        keep = [l for l in det.splitlines() if any(s in l for s in ('Duration', 'Theoretical Occupancy', 'Achieved Occupancy',
                'Registers Per', 'Shared Memory', 'Compute (SM) Throughput', 'Memory Throughput', 'L2 Cache Throughput',
                'Executed Ipc', 'Issue Slots Busy', 'No Eligible', 'Stall', 'OPT', 'Est. Speedup'))]
        f.write('\n## selected lines of --page details\n' + '\n'.join(keep[:80]) + '\n')


for name in sorted(os.listdir(src)):
    p = os.path.join(src, name)
    if name.startswith(tag) and name.endswith('launches.csv'):
        launches(p, os.path.join(dst, name.replace('.csv', '.txt')))
    elif name.startswith(tag) and name.endswith('.ncu-rep'):
        full(p, os.path.join(dst, name.replace('.ncu-rep', '_ncu_full.txt')))
    elif name.startswith(tag) and (name.endswith('_bench.log') or name.endswith('_bench_ref.log') or name.endswith('probe.log')
                                   or name.endswith('tests.log')):
        open(os.path.join(dst, name.replace('.log', '.txt')), 'w').write(open(p).read())
if os.path.exists(os.path.join(src, 'parity_report.txt')):
    open(os.path.join(dst, f'{tag}_parity_report.txt'), 'w').write(open(os.path.join(src, 'parity_report.txt')).read())
print(sorted(os.listdir(dst)))

"""Summarise an `ncu --set full` report (.ncu-rep) into the handful of numbers the design notes
quote: duration, DRAM bytes, L2 bytes, tensor-pipe activity, shared-memory wavefronts, registers,
and the top stall sites.  Usage: python profiles/summarize_ncu.py report.ncu-rep [n_stall_rows]
       [--traffic-json out.json --rows N]   (also write {rows, dram_bytes} of the captured launch: bench.py's
                                             `roofline.traffic` reads profiles/suffstats_traffic.json)"""
import csv
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed.sum',
        'sm__cycles_elapsed.max', 'smsp__cycles_active.avg', 'launch__registers_per_thread',
        'launch__grid_size', 'launch__block_size', 'launch__shared_mem_per_block_dynamic',
        'sm__warps_active.avg.pct_of_peak_sustained_active']


UNIT_SCALE = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'Tbyte': 1e12}


def main():
    rep = sys.argv[1]
    argv = sys.argv[2:]
    traffic_json = rows_arg = None
    if '--traffic-json' in argv:
        i = argv.index('--traffic-json')
        traffic_json = argv[i + 1]
        del argv[i:i + 2]
    if '--rows' in argv:
        i = argv.index('--rows')
        rows_arg = int(argv[i + 1])
        del argv[i:i + 2]
    n_rows = int(argv[0]) if argv else 12
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print('report: %s' % rep)
    print('kernel: %s' % vals[hdr.index('Kernel Name')][:100])
    for key in KEYS:
        if key in hdr:
            i = hdr.index(key)
            print('  %-64s %s %s' % (key, vals[i], units[i]))
    if traffic_json:
        import json
        total_bytes = 0.0
        for key in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            i = hdr.index(key)
            total_bytes += float(vals[i].replace(',', '')) * UNIT_SCALE[units[i]]
        with open(traffic_json, 'w') as fh:
            json.dump({'rows': rows_arg, 'dram_bytes': int(total_bytes), 'report': rep,
                       'kernel': vals[hdr.index('Kernel Name')][:60]}, fh)
            fh.write('\n')
    src = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    h, data = rows[1], rows[2:]
    ia, isrc, ins, iex = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
    stall_cols = [i for i, c in enumerate(h) if c.startswith('stall_') and 'Not Issued' not in c]
    total = sum(int(r[ins] or 0) for r in data)
    print('  warp-stall samples: %d over %d SASS instructions; top sites:' % (total, len(data)))
    for r in sorted(data, key=lambda r: -int(r[ins] or 0))[:n_rows]:
        st = sorted(((int(r[i] or 0), h[i][6:]) for i in stall_cols), reverse=True)[:2]
        print('    %6s samples  %10s exec  %-56s %s' % (r[ins], r[iex], r[isrc][:56], [s for s in st if s[0]]))
    reasons = {}
    for r in data:
        for i in stall_cols:
            reasons[h[i][6:]] = reasons.get(h[i][6:], 0) + int(r[i] or 0)
    top = sorted(reasons.items(), key=lambda kv: -kv[1])[:6]
    print('  stall reasons (all warps): ' + ', '.join('%s %.0f%%' % (k, 100.0 * v / max(total, 1)) for k, v in top))


if __name__ == '__main__':
    main()

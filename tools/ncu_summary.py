"""Summarise an ncu report (.ncu-rep from `ncu --set full`) as text: per captured launch the duration, DRAM bytes, tensor-pipe
activity, occupancy, registers and the top warp-stall sites (needs -lineinfo / --import-source on).
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/rNN_x_summary.txt"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'smsp__inst_executed.sum', 'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__cycles_elapsed.avg']


def run(args):
    return subprocess.run(['ncu'] + args, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout


def main():
    rep = sys.argv[1]
    raw = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = raw[0], raw[1]
    ix = {h: i for i, h in enumerate(hdr)}
    print('# %s' % rep)
    for li, r in enumerate(raw[2:]):
        print('\n== launch %d: %s  grid %s block %s' % (li, r[ix['Kernel Name']][:100], r[ix.get('Grid Size', 0)], r[ix.get('Block Size', 0)]))
        for k in KEYS:
            if k in ix:
                print('   %-72s %s %s' % (k, r[ix[k]], units[ix[k]]))
        src = list(csv.reader(io.StringIO(run(['-i', rep, '--page', 'source', '--csv', '--launch-skip', str(li), '--launch-count', '1']))))
        if len(src) < 3:
            continue
        sh = src[1]
        sx = {h: i for i, h in enumerate(sh)}
        if '# Samples' not in sx:
            continue
        seen, rows = set(), []
        for s in src[2:]:
            if len(s) < len(sh) or s[sx['Address']] in seen:
                continue
            seen.add(s[sx['Address']])
            try:
                n = int(s[sx['# Samples']])
            except ValueError:
                continue
            rows.append((n, s))
        tot = sum(n for n, _ in rows) or 1
        stalls = [h for h in sh if h.startswith('stall_') and 'Not Issued' not in h]
        agg = {}
        for n, s in rows:
            for st in stalls:
                try:
                    agg[st] = agg.get(st, 0) + int(s[sx[st]] or 0)
                except ValueError:
                    pass
        print('   warp-stall samples: %d; by reason: %s' % (tot, ', '.join('%s %.0f%%' % (k[6:], 100.0 * v / tot) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6])))
        print('   top stall sites (samples, executions, SASS):')
        for n, s in sorted(rows, key=lambda t: -t[0])[:10]:
            print('     %6d %10s  %s' % (n, s[sx['Instructions Executed']], s[sx['Source']][:90]))


if __name__ == '__main__':
    main()

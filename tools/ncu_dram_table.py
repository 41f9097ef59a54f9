"""ncu report of tools/profile_step.py -> profiles/rNN_ncu_dram_bytes.json: measured DRAM bytes (read + write) per
sample (per env for the preprocessing kernel) of every hot-path kernel, keyed by bench.py's kernel names.
usage: python tools/ncu_dram_table.py gpurun_out/x.ncu-rep BATCH ENVS > profiles/r01_ncu_dram_bytes.json"""
import csv, io, json, re, subprocess, sys

NAMES = [(r'conv1_i8_kernel', 'conv1_fwd'), (r'convk_kernel<1>', 'conv2_fwd'), (r'convk_kernel<2>', 'conv3_fwd'), (r'convk_kernel<7>', 'conv3_fwd'),
         (r'convk_kernel<3>', 'conv3_dgrad'), (r'convk_kernel<4>', 'conv2_dgrad'), (r'stream_gemm_kernel<128, 0>', 'fc4_fwd'),
         (r'stream_gemm_kernel<128, 1>', 'fc4_dgrad'), (r'stream_gemm_kernel<128, 2>', 'fc4_wgrad'),
         (r'wgrad2_kernel<0>', 'conv1_wgrad'), (r'wgrad2_kernel<1>', 'conv2_wgrad'), (r'wgrad2_kernel<2>', 'conv3_wgrad'),
         (r'preprocess_u8_pipe_kernel', 'preprocess_u8'), (r'preprocess_u8_kernel', 'preprocess_u8'), (r'heads_fwd', 'heads_fwd'), (r'heads_bwd', 'heads_bwd')]


def gb(value, unit):
    v = float(value.replace(',', ''))
    return v * {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


def main():
    rep, batch, envs = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    table = {}
    for r in rows[2:]:
        name = r[ix['Kernel Name']]
        for pat, key in NAMES:
            if pat in name:
                rd = gb(r[ix['dram__bytes_read.sum']], units[ix['dram__bytes_read.sum']])
                wr = gb(r[ix['dram__bytes_write.sum']], units[ix['dram__bytes_write.sum']])
                n = envs if key == 'preprocess_u8' else batch
                table[key] = {'dram_bytes_per_unit': (rd + wr) / n, 'read': rd / n, 'write': wr / n,
                              'duration_us': float(r[ix['gpu__time_duration.sum']].replace(',', '')) * {'us': 1, 'ms': 1e3, 'ns': 1e-3}.get(units[ix['gpu__time_duration.sum']], 1),
                              'tensor_pipe_active_pct': float(r[ix['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active']])}
    print(json.dumps({'source': rep.split('/')[-1], 'batch': batch, 'envs': envs, 'kernels': table}, indent=1))


main()

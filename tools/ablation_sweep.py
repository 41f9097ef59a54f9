"""What is each tensor-core kernel waiting for?  Run bench.py once per PAACB_DBG ablation switch (a stage of a kernel family
is switched off; results are wrong, only the timing matters) and tabulate the per-kernel device time per step.

    python tools/ablation_sweep.py > gpurun_out/r01_ablations.json          (one B200, about 3 minutes)

Switches (csrc/tc2_conv1.cu, tc2_wgrad.cu, tc2_conv.cu): 1 / 2 / 4 / 8 conv1 forward: no stores / no epilogue arithmetic /
one MMA of eight / no A loads; 16 / 32 / 64 / 128 weight gradients: hi*hi MMAs only / no TMA loads / no uint8 conversion /
half the k-tiles; 256 / 512 / 1024 conv2-conv3 forward and data gradients: no A_lo MMAs / no stores / no patch loads.
"""
import json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SWITCHES = [(0, 'baseline'), (1, 'conv1 fwd: no stores'), (2, 'conv1 fwd: no epilogue arithmetic'), (4, 'conv1 fwd: 1 of 8 MMAs'),
            (8, 'conv1 fwd: no A loads'), (16, 'wgrad: hi*hi MMAs only'), (32, 'wgrad: no TMA loads'),
            (64, 'conv1 wgrad: no uint8 conversion'), (128, 'wgrad: half the k-tiles'),
            (256, 'conv fwd/dgrad: no A_lo MMAs'), (512, 'conv fwd/dgrad: no stores'), (1024, 'conv fwd/dgrad: no patch loads'),
            (2048, 'fc fwd/dgrad/wgrad: no TMA loads'), (16384, 'conv dgrad: no ReLU-mask loads')]
KERNELS = ['conv1_fwd', 'conv2_fwd', 'conv3_fwd', 'fc4_fwd', 'conv2_dgrad', 'conv3_dgrad', 'fc4_dgrad', 'conv1_wgrad', 'conv2_wgrad',
           'conv3_wgrad', 'fc4_wgrad']


def main():
    # the switches exist only in the measurement build (PAACB_ABLATIONS=1 python -m paac_b200.build -> libpaacb_abl.so)
    abl = os.path.join(ROOT, 'paac_b200', 'libpaacb_abl.so')
    if not os.path.exists(abl):
        raise SystemExit('build the measurement library first: PAACB_ABLATIONS=1 python -m paac_b200.build')
    only = [int(a) for a in sys.argv[1:]]           # optional: the switches to run (0 = baseline is always run)
    rows = []
    for dbg, what in SWITCHES:
        if only and dbg and dbg not in only:
            continue
        env = dict(os.environ, PAACB_DBG=str(dbg), PAACB_LIB=abl)
        out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '10', '--no_cpu_baseline', '--no_variants',
                              '--no_e2e'], env=env, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
        line = [l for l in out.splitlines() if l.startswith('{')][-1]
        d = json.loads(line)
        ks = {k['name']: k['ms'] / d['steps'] for k in d['kernels']}
        rows.append({'PAACB_DBG': dbg, 'switched_off': what, 'sm_mhz': d['clocks']['sm_mhz'],
                     'ms_per_step': {k: round(ks[k], 4) for k in KERNELS}})
    base = rows[0]['ms_per_step']
    for r in rows[1:]:
        r['change_vs_baseline'] = {k: round(r['ms_per_step'][k] / base[k] - 1.0, 3) for k in KERNELS
                                   if abs(r['ms_per_step'][k] / base[k] - 1.0) >= 0.02}
    print(json.dumps({'what': 'per-kernel device time (ms per update cycle of 20,480 env-steps, NatureNetwork, bf16x3) with one '
                              'stage of a kernel family switched off; the results of an ablated run are wrong, only the '
                              'timing is meaningful',
                      'caveat': 'under sw_power_cap, switching work off anywhere raises the clocks for everything else: kernels outside '
                                'the ablated family move by a uniform few per cent; read the large changes in the targeted family',
                      'rows': rows}, indent=1))


if __name__ == '__main__':
    main()

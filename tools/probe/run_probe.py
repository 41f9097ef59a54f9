"""Drive tools/probe/libprobe.so on a B200: read out, for a list of operand-descriptor variants, which shared-memory
element the tensor core fetches for every (row, k), and compare with candidate addressing models.

    python tools/probe/run_probe.py            # all cases, each in its own process -> gpurun_out/probe.json
"""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
IMG = 160 * 1024
OFF_ID = 128 * 1024          # where the identity operand lives (K-major SW128, 1024-aligned)


def bf16_bits(x):
    return (np.asarray(x, np.float32).view(np.uint32) >> 16).astype(np.uint16)


def desc(addr, lbo, sbo, layout, base_offset=0):
    return ((addr >> 4) & 0x3FFF) | (((lbo >> 4) & 0x3FFF) << 16) | (((sbo >> 4) & 0x3FFF) << 32) | (1 << 46) | \
        ((base_offset & 7) << 49) | (layout << 61)


def idesc(a_major, b_major, n=32, m=128):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_major << 15) | (b_major << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def identity_image(img16, rows):
    """identity operand [rows][K=16], K-major SW128 at OFF_ID: I[r][k] = (r == k)"""
    one = bf16_bits(1.0)
    for r in range(min(rows, 16)):
        k = r
        byte = OFF_ID + r * 128 + (((k // 8) ^ (r % 8)) * 16) + (k % 8) * 2
        img16[byte // 2] = one


def run_lib(image, adesc, bdesc, idsc, nmma=1, a_step=0, b_step=0):
    lib = ctypes.CDLL(os.path.join(HERE, 'libprobe.so'))
    lib.probe_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint32,
                              ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32, ctypes.c_void_p]
    out = np.zeros((128, 32), np.float32)
    rc = lib.probe_run(image.ctypes.data, image.nbytes, adesc, bdesc, idsc, nmma, a_step, b_step, out.ctypes.data)
    if rc != 0:
        raise RuntimeError('probe_run rc=%d' % rc)
    return out


def readout(which, d_test, major):
    """element index (in 2-byte units from the image base) fetched for each (row, k) of the operand under test.
    which = 'A': returns [128][16]; which = 'B': returns [32][16]."""
    res = []
    for part in (0, 1):
        img16 = np.zeros(IMG // 2, np.uint16)
        n_el = OFF_ID // 2
        idx = np.arange(n_el)
        vals = (idx & 0xFF) if part == 0 else (idx >> 8)
        img16[:n_el] = bf16_bits(vals.astype(np.float32))
        if which == 'A':
            identity_image(img16, 32)
            out = run_lib(img16, d_test, desc(OFF_ID, 16, 1024, 2), idesc(major, 0))
            res.append(out[:, :16])                 # D[m][n] = A[m][k = n]
        else:
            identity_image(img16, 128)
            out = run_lib(img16, desc(OFF_ID, 16, 1024, 2), d_test, idesc(0, major))
            res.append(out[:16, :].T)               # D[m][n] = B[n][k = m]
    return (res[0] + 256.0 * res[1]).astype(np.int64)


# ---- addressing models: byte offset of element (r, k) -------------------------------------------------
def model_kmajor_abs(start, sbo, row_bytes, r, k):
    """K-major, rows `row_bytes` apart inside an 8-row group, swizzle XOR on absolute address bits"""
    lin = start + (r // 8) * sbo + (r % 8) * row_bytes + k * 2
    nb = {128: 3, 64: 2, 32: 1}[row_bytes]         # Swizzle<nb, 4, 3>: address bits [7, 7+nb) XORed into bits [4, 4+nb)
    return lin ^ (((lin >> 7) & ((1 << nb) - 1)) << 4)


def model_mn_abs(start, lbo, sbo, r, k):
    """MN-major SW128 (bf16): 64 MN elements = one 128-byte row; K index = row"""
    lin = start + (r // 64) * lbo + (k // 8) * sbo + (k % 8) * 128 + (r % 64) * 2
    return lin ^ (((lin >> 7) & 7) << 4)


CASES = {}


def case(name):
    def deco(f):
        CASES[name] = f
        return f
    return deco


def summarize(got, want):
    got = np.asarray(got)
    want = np.asarray(want) // 2
    ok = bool(np.array_equal(got, want))
    bad = int((got != want).sum())
    return {'match': ok, 'mismatches': bad}


def k_case(which, start, sbo, layout, row_bytes, base_offset=0, rows=None):
    d = desc(start, 16, sbo, layout, base_offset)
    got = readout(which, d, 0)
    rows = got.shape[0]
    want = np.array([[model_kmajor_abs(start, sbo, row_bytes, r, k) for k in range(16)] for r in range(rows)])
    s = summarize(got, want)
    s['sample'] = (got[:10, :] * 2).tolist()        # byte offsets of the first rows
    s['sample_r64'] = (got[64:66, :] * 2).tolist() if rows > 64 else None
    return s


def mn_case(which, start, lbo, sbo, base_offset=0):
    d = desc(start, lbo, sbo, 2, base_offset)
    got = readout(which, d, 1)
    rows = got.shape[0]
    want = np.array([[model_mn_abs(start, lbo, sbo, r, k) for k in range(16)] for r in range(rows)])
    s = summarize(got, want)
    s['sample_r0_8'] = (got[:9, :] * 2).tolist()
    s['sample_r64'] = (got[64:66, :] * 2).tolist() if rows > 64 else None
    return s


for _j in (0, 1, 2, 3, 8, 9, 10, 13):
    case('A_k128_start%d' % (_j * 128))(lambda j=_j: k_case('A', j * 128, 1024, 2, 128))
for _j in (1, 2, 10):
    case('A_k128_start%d_bo' % (_j * 128))(lambda j=_j: k_case('A', j * 128, 1024, 2, 128, base_offset=j % 8))
for _o in (32, 64, 96, 128 * 3 + 64, 1280 + 32):
    case('A_k128_start%d' % _o)(lambda o=_o: k_case('A', o, 1024, 2, 128))
for _sbo in (1152, 1280, 2048, 2560):
    case('A_k128_sbo%d' % _sbo)(lambda s=_sbo: k_case('A', 0, s, 2, 128))
for _j in (0, 1, 3, 10):
    case('B_k128_start%d' % (_j * 128))(lambda j=_j: k_case('B', j * 128, 1024, 2, 128))
for _o in (0, 32, 64, 96, 256, 32 * 21):
    case('A_k32_start%d' % _o)(lambda o=_o: k_case('A', o, 256, 6, 32))
for _sbo in (512, 672):
    case('A_k32_sbo%d' % _sbo)(lambda s=_sbo: k_case('A', 0, s, 6, 32))
for _o in (0, 64, 128, 640):
    case('A_k64_start%d' % _o)(lambda o=_o: k_case('A', o, 512, 4, 64))
for _lbo, _sbo, _st in ((8192, 1024, 0), (128, 1024, 0), (256, 1024, 0), (8192, 1024, 128), (8192, 1024, 1280),
                        (128, 1024, 1280), (16384, 2048, 0), (1024, 8192, 0)):
    case('A_mn_lbo%d_sbo%d_start%d' % (_lbo, _sbo, _st))(lambda l=_lbo, s=_sbo, t=_st: mn_case('A', t, l, s))
for _st in (0, 128, 1280):
    case('B_mn_start%d' % _st)(lambda t=_st: mn_case('B', t, 8192, 1024))
case('A_mn_start1280_bo')(lambda: mn_case('A', 1280, 8192, 1024, base_offset=(1280 >> 7) & 7))


def main():
    if len(sys.argv) > 2 and sys.argv[1] == '--case':
        print('RESULT ' + json.dumps(CASES[sys.argv[2]]()))
        return
    results = {}
    for name in CASES:
        try:
            r = subprocess.run([sys.executable, __file__, '--case', name], capture_output=True, text=True, timeout=120)
            line = [l for l in r.stdout.splitlines() if l.startswith('RESULT ')]
            results[name] = json.loads(line[0][7:]) if line else {'error': (r.stdout + r.stderr)[-400:]}
        except subprocess.TimeoutExpired:
            results[name] = {'error': 'timeout'}
        print(name, {k: v for k, v in results[name].items() if k in ('match', 'mismatches', 'error')}, flush=True)
    os.makedirs('gpurun_out', exist_ok=True)
    with open('gpurun_out/probe.json', 'w') as f:
        json.dump(results, f)


if __name__ == '__main__':
    main()

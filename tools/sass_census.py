#!/usr/bin/env python
"""Per-kernel SASS opcode census of the shipped library: which kernels carry tcgen05 / TMA / TMEM instructions.

    python tools/sass_census.py [paac_b200/libpaacb.so] > profiles/r02_sass_census.txt

Runs `cuobjdump -sass` (no GPU needed) and counts, per kernel, the opcodes that prove a Blackwell-native path
(/opt/skills/guides/B200_PROFILING.md: tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG/UBLKCP,
TMA prefetch -> UTMAPF; HMMA would be a legacy mma.sync path), plus the wide global accesses and reductions the DESIGN
notes refer to (STG...256, REDG).  The .so itself is git-ignored (it travels to the GPU box by snapshot), so this file is
the committed record of what was built.
"""
import collections
import hashlib
import re
import subprocess
import sys

OPS = [('UTC*MMA', re.compile(r'\bUTC[A-Z]*MMA\b')), ('UTMALDG', re.compile(r'\bUTMALDG\b')), ('UTMASTG', re.compile(r'\bUTMASTG\b')),
       ('UTMAPF', re.compile(r'\bUTMAPF\b')), ('UBLKCP', re.compile(r'\bUBLKCP\b')), ('LDTM', re.compile(r'\bLDTM\b')),
       ('STTM', re.compile(r'\bSTTM\b')), ('HMMA', re.compile(r'\bHMMA\b')), ('STG.256', re.compile(r'\bSTG\.[A-Z0-9.]*256\b')),
       ('LDG.256', re.compile(r'\bLDG\.[A-Z0-9.]*256\b')), ('REDG', re.compile(r'\bREDG\b')), ('LDGSTS', re.compile(r'\bLDGSTS\b'))]
KIND = re.compile(r'\b(UTC[A-Z]*MMA)\b')


def demangle(names):
    try:
        out = subprocess.run(['c++filt'], input='\n'.join(names), capture_output=True, text=True, check=True).stdout.split('\n')
        return dict(zip(names, out))
    except Exception:
        return {n: n for n in names}


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else 'paac_b200/libpaacb.so'
    sass = subprocess.run(['cuobjdump', '-sass', lib], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r'arch = (sm_\w+)', sass)))
    counts, kinds, order, fn = collections.defaultdict(collections.Counter), collections.defaultdict(set), [], None
    for line in sass.split('\n'):
        m = re.search(r'Function : (\S+)', line)
        if m:
            fn = m.group(1)
            order.append(fn)
            continue
        if fn is None or '/*' not in line:
            continue
        counts[fn]['instructions'] += 1
        for name, rx in OPS:
            if rx.search(line):
                counts[fn][name] += 1
        k = KIND.search(line)
        if k:
            kinds[fn].add(k.group(1))
    names = demangle(order)
    sha = hashlib.sha256(open(lib, 'rb').read()).hexdigest()[:16]
    print('# SASS opcode census of %s (sha256 %s...), cubin archs: %s' % (lib, sha, ', '.join(archs)))
    print('# columns: kernel | instructions | ' + ' | '.join(n for n, _ in OPS) + ' | MMA opcodes')
    for fn in order:
        c = counts[fn]
        short = re.sub(r'^void ', '', names[fn])
        short = re.sub(r'\(.*\)$', '', short)
        print('%-64s %6d  ' % (short[:64], c['instructions']) + ' '.join('%5d' % c[n] for n, _ in OPS) + '  ' + ','.join(sorted(kinds[fn])))
    tc = [fn for fn in order if counts[fn]['UTC*MMA']]
    print('# %d kernels, %d with tcgen05.mma (UTC*MMA), %d with TMA tensor loads (UTMALDG), %d with HMMA (legacy mma.sync)' % (
        len(order), len(tc), sum(1 for fn in order if counts[fn]['UTMALDG']), sum(1 for fn in order if counts[fn]['HMMA'])))


if __name__ == '__main__':
    main()

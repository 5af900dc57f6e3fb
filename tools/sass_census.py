"""Instruction census of libbyolo.so per kernel (cuobjdump -sass): the mnemonics that prove the tcgen05 / TMA / TMEM
path (UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA load / store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,
UTCATOMSWS/... = TMEM alloc), the legacy HMMA (mma.sync, stem only) and totals.  Output: profiles/r02/sass_census.txt."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'bayesian-yolov3_b200', 'byolo', 'libbyolo.so')
KEYS = ['UTCHMMA', 'UTCHMMA.2CTA', 'UTMALDG', 'UTMALDG.*IM2COL', 'UTMASTG', 'LDTM', 'UTCBAR', 'UTCBAR.*MULTICAST', 'UTCATOMSWS', 'SYNCS',
        'HMMA', 'FFMA', 'F2FP', 'IMAD.WIDE.U32|IMAD.HI', 'STS', 'LDS', 'LDG', 'STG', 'MUFU', 'BAR.SYNC', 'UCGABAR']


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(['c++filt', n], stdout=subprocess.PIPE, text=True).stdout.strip()
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = demangle(m.group(1))
            kernels[cur] = []
            continue
        m = re.match(r'\s+/\*[0-9a-f]+\*/\s+(.*?);', line)
        if m and cur is not None:
            kernels[cur].append(m.group(1))
    out = ['SASS instruction census of bayesian-yolov3_b200/byolo/libbyolo.so (cuobjdump -sass, sm_100a), per kernel', '']
    for name, ins in kernels.items():
        short = re.sub(r'\(.*', '', name)
        counts = [(k, sum(1 for i in ins if re.search(r'(^|\s)(@!?U?P\d+\s+)?(%s)' % k, i))) for k in KEYS]
        out.append('%s   [%d instructions]' % (short, len(ins)))
        out.append('    ' + '  '.join('%s=%d' % kc for kc in counts if kc[1]))
    text = '\n'.join(out) + '\n'
    dst = os.path.join(ROOT, 'profiles', 'r02', 'sass_census.txt')
    open(dst, 'w').write(text)
    sys.stdout.write(text)


if __name__ == '__main__':
    main()

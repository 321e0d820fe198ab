"""Histogram of the Blackwell-specific SASS opcodes per kernel of libcoldbrew_b200.so (cuobjdump -sass):
UTCHMMA / UTCQMMA (tcgen05.mma), UTMALDG (TMA tensor loads), UBLKPF (bulk L2 prefetch), LDTM (tcgen05.ld),
UTCBAR (tcgen05.commit), SYNCS (mbarrier).  Written to profiles/sass_opcodes.txt.

    python scripts/sass_opcodes.py
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'gnn_tail_generalization_b200', 'libcoldbrew_b200.so')
WATCH = ('UTCHMMA', 'UTCQMMA', 'UTCOMMA', 'UTMALDG', 'UTMAPF', 'UBLKPF', 'UBLKCP', 'LDTM', 'STTM', 'UTCBAR', 'UTCATOMSWS',
         'SYNCS', 'LDG', 'STG', 'SHFL', 'FADD', 'FMUL', 'FFMA', 'HMMA')


def main():
    sass = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r'arch = (sm_\w+)', sass)))
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r'\s*Function : (\S+)', line)
        if m:
            cur = subprocess.run(['c++filt', m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r'\(.*', '', cur)
            per[cur] = collections.Counter()
            continue
        m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)', line)
        if m and cur is not None:
            op = m.group(1)
            for w in WATCH:
                if op.startswith(w):
                    per[cur][w] += 1
                    break
    out = [f'# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: cubin architectures {arch}',
           '# opcode counts per kernel (static instruction counts; only the watched opcode families are listed)', '']
    for k, c in per.items():
        if not c:
            continue
        out.append(k)
        out.append('    ' + '  '.join(f'{op}={n}' for op, n in sorted(c.items()) if n))
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    out += ['', 'TOTAL  ' + '  '.join(f'{op}={n}' for op, n in sorted(tot.items()))]
    path = os.path.join(ROOT, 'profiles', 'sass_opcodes.txt')
    open(path, 'w').write('\n'.join(out) + '\n')
    print(path, {k: tot[k] for k in ('UTCHMMA', 'UTMALDG', 'LDTM', 'UBLKPF', 'UTCBAR')})


if __name__ == '__main__':
    sys.exit(main())

"""Per-kernel counts of the tensor-core / TMEM / TMA instructions in the shipped library (cuobjdump -sass; no GPU needed):
    python tools/sass_summary.py > profiles/r02_sass_summary.txt
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store,
UTCBAR = tcgen05.commit, HMMA = mma.sync (none expected).  Mnemonics as in /opt/skills/guides/B200_PROFILING.md."""
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'conformer_ocr_b200', 'libcocr_b200.so')
MNEMONICS = ('UTCHMMA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UTCBAR', 'HMMA', 'FFMA2')


def main():
    tool = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    sass = subprocess.run([tool, '-sass', LIB], capture_output=True, text=True, check=True).stdout
    demangle = shutil.which('c++filt')
    rows = []
    for part in re.split(r'\n\s*Function : ', sass)[1:]:
        name = part.split('\n', 1)[0].strip()
        c = {m: len(re.findall(r'\b' + m, part)) for m in MNEMONICS}
        n_instr = len(re.findall(r'^\s+/\*[0-9a-f]{4,6}\*/\s+\S', part, flags=re.M))
        rows.append((name, n_instr, c))
    if demangle:
        out = subprocess.run([demangle], input='\n'.join(r[0] for r in rows), capture_output=True, text=True).stdout.split('\n')
        rows = [(re.sub(r'\(.*', '', d.replace('(anonymous namespace)::', '')) or r[0], r[1], r[2]) for d, r in zip(out, rows)]
    print(f'# {os.path.relpath(LIB, ROOT)}: {len(rows)} kernels; instruction counts per kernel (cuobjdump -sass)')
    print(f'{"kernel":64s} {"instr":>7s} ' + ' '.join(f'{m:>8s}' for m in MNEMONICS))
    for name, n, c in sorted(rows, key=lambda r: (-r[2]['UTCHMMA'], r[0])):
        print(f'{name[:64]:64s} {n:7d} ' + ' '.join(f'{c[m]:8d}' for m in MNEMONICS))
    return 0


if __name__ == '__main__':
    sys.exit(main())

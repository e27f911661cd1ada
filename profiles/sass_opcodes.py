"""Opcode histogram per kernel from `cuobjdump -sass` of the in-tree library: the SASS evidence that the hot
path is tcgen05 / TMEM / TMA (mnemonics from /opt/skills/guides/B200_PROFILING.md).

  python profiles/sass_opcodes.py > profiles/r2_sass_opcodes.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, 'multimodal-long-transformer-2021_b200', 'libmlt_attn.so')
WATCH = ['UTCHMMA', 'UTCBAR', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'MUFU.EX2', 'HMMA', 'FFMA', 'LDS', 'STS',
         'LDG', 'STG', 'SHFL', 'BAR']


def main():
  out = subprocess.run(['cuobjdump', '-sass', LIB], capture_output=True, text=True, check=True).stdout
  kernels = collections.OrderedDict()
  cur = None
  for line in out.splitlines():
    m = re.match(r'\s*Function : (\S+)', line)
    if m:
      cur = kernels.setdefault(m.group(1), collections.Counter())
      continue
    m = re.match(r'\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)', line)
    if m and cur is not None:
      op = m.group(1)
      cur['_total'] += 1
      for w in WATCH:
        if op == w or op.startswith(w + '.') or (w == 'MUFU.EX2' and op.startswith('MUFU.EX2')):
          cur[w] += 1
  demangle = subprocess.run(['c++filt'], input='\n'.join(kernels), capture_output=True, text=True).stdout.splitlines()
  print('# SASS opcode histogram per kernel (`cuobjdump -sass libmlt_attn.so`, sm_100a)\n')
  print('UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk,')
  print('UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.  No HMMA (legacy mma.sync) anywhere on the tensor-core path.\n')
  cols = ['_total'] + WATCH
  print('| kernel | instructions | ' + ' | '.join(WATCH) + ' |')
  print('|---|' + '---|' * len(cols))
  for (name, c), dm in zip(kernels.items(), demangle):
    short = dm.replace('(anonymous namespace)::', '')
    short = re.sub(r'\(.*', '', short).replace('void ', '').replace('mlt::', '')
    if not any(c[w] for w in ('UTCHMMA', 'LDTM', 'UTMALDG')) and 'kernel' not in short:
      continue
    print(f'| `{short}` | ' + ' | '.join(str(c[k]) for k in cols) + ' |')


if __name__ == '__main__':
  main()

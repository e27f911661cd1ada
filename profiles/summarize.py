"""Turns the ncu artefacts a gpurun call brought back (gpurun_out/) into the committed summaries.

  python profiles/summarize.py gpurun_out/r1_launches.csv gpurun_out/r1_prof_all.ncu-rep r1

writes profiles/<tag>_launches.md (per-kernel share of the step from the cold-cache, serialised
`--metrics gpu__time_duration.sum` pass) and profiles/<tag>_ncu_full.md (`--set full` metrics).
"""
import collections
import csv
import io
import subprocess
import sys


def launches(path):
  rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
  hdr = rows[0]
  ik, iv, im = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Name')
  agg = collections.OrderedDict()
  for r in rows[1:]:
    if r[im] != 'gpu__time_duration.sum':
      continue
    name = r[ik].split('(')[0].split('::')[-1]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(',', ''))
  return agg


def full(rep):
  raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
  rows = list(csv.reader(io.StringIO(raw)))
  hdr = rows[0]
  keys = ['launch__grid_size', 'launch__registers_per_thread', 'gpu__time_duration.sum',
          'dram__bytes_read.sum', 'dram__bytes_write.sum',
          'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
          'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
          'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
          'smsp__issue_active.avg.pct_of_peak_sustained_active',
          'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
          'sm__icc_request_hit_rate.pct',
          'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
          'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio']
  units = dict(zip(hdr, rows[1]))
  out = []
  for r in rows[2:]:
    d = dict(zip(hdr, r))
    out.append((d['Kernel Name'].split('(')[0].split('::')[-1], {k: (d.get(k), units.get(k)) for k in keys}))
  return out


def main():
  lpath, rep, tag = sys.argv[1], sys.argv[2], sys.argv[3]
  agg = launches(lpath)
  tot = sum(v[1] for v in agg.values())
  with open(f'profiles/{tag}_launches.md', 'w') as f:
    f.write(f'# {tag}: launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline`\n\n'
            '`ncu --metrics gpu__time_duration.sum --clock-control none -c 200` (cold-cache, serialised:\n'
            'compare SHARES, not absolutes).  Time unit as reported by ncu (ns).\n\n'
            '| kernel | launches | total | share |\n|---|---|---|---|\n')
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
      f.write(f'| `{k}` | {n} | {t:,.0f} | {100 * t / tot:.1f}% |\n')
  with open(f'profiles/{tag}_ncu_full.md', 'w') as f:
    f.write(f'# {tag}: `ncu --set full --clock-control none -k regex:tc_|gl2_ -c 7` on `bench.py --steps 1 --warmup 3 --no-cpu-baseline` '
            '(workload c3_4096)\n\n')
    for name, d in full(rep):
      f.write(f'## `{name}` grid {d["launch__grid_size"][0]}\n\n| metric | value | unit |\n|---|---|---|\n')
      for k, (v, u) in d.items():
        f.write(f'| {k} | {v} | {u} |\n')
      f.write('\n')
  # DRAM traffic per launch of every captured kernel, keyed by the name bench.py's per-kernel
  # timing uses (capture order of one step: fwd global/long rows, bwd_q global/long, bwd_kv long/global)
  import json
  traffic = {}
  for name, d in full(rep):
    grid = int(d['launch__grid_size'][0])
    def as_bytes(key):
      v, u = d[key]
      scale = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(u, 1)
      return float(v) * scale
    full_name = name
    if 'gl2_fwd_long' in full_name:
      key = 'gl2_fwd_long_rows'
    elif 'gl2_bwd_q_long' in full_name:
      key = 'gl2_bwd_q_long_rows'
    elif 'tc_fwd_kernel' in full_name:
      key = 'fwd'
    elif 'tc_bwd_q_kernel' in full_name:
      key = 'bwd_q'
    elif 'tc_bwd_kv_kernel' in full_name:
      key = 'bwd_kv'
    else:
      continue
    traffic.setdefault(key, []).append((grid, as_bytes('dram__bytes_read.sum') + as_bytes('dram__bytes_write.sum')))
  out = {}
  for kind, lst in traffic.items():
    lst.sort()
    small, large = lst[0], lst[-1]
    if kind.startswith('gl2_'):     # persistent kernels: one launch per step, named like bench.py's timing key
      out[kind] = {'grid': large[0], 'dram_bytes_per_launch': large[1]}
      continue
    names = {'fwd': ('tc_fwd_global_rows', 'tc_fwd_long_rows'),
             'bwd_q': ('tc_bwd_q_global_rows', 'tc_bwd_q_long_rows'),
             'bwd_kv': ('tc_bwd_kv_global_keys', 'tc_bwd_kv_long_keys')}[kind]
    if len(lst) > 1 and small[0] != large[0]:
      out[names[0]] = {'grid': small[0], 'dram_bytes_per_launch': small[1]}
      out[names[1]] = {'grid': large[0], 'dram_bytes_per_launch': large[1]}
    else:                            # only the global-row launch of this family is left on the general kernels
      out[names[0]] = {'grid': small[0], 'dram_bytes_per_launch': small[1]}
  out['_source'] = f'{tag}: ncu --set full --clock-control none -k regex:tc_|gl2_ on bench.py --steps 1 --warmup 3 (workload c3_4096, batch 16)'
  json.dump(out, open('profiles/traffic.json', 'w'), indent=1)
  print('wrote', f'profiles/{tag}_launches.md', f'profiles/{tag}_ncu_full.md', 'profiles/traffic.json')


if __name__ == '__main__':
  main()

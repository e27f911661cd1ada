"""PCIe copy bandwidth of the box: pinned host <-> device, each direction alone and both at once.

The end-to-end line of bench.py is bound by these copies (428 MB each way per step); this probe
says what the box can deliver so that the pipelining can be judged against it.
"""
import time
import torch

MB = 1 << 20
n = 428 * MB
dev = torch.device('cuda', 0)
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(h2d, d2h, reps=10, pieces=1):
  torch.cuda.synchronize()
  t0 = time.perf_counter()
  step = n // pieces
  for _ in range(reps):
    if h2d:
      with torch.cuda.stream(s1):
        for k in range(pieces):
          d_in[k * step:(k + 1) * step].copy_(h_in[k * step:(k + 1) * step], non_blocking=True)
    if d2h:
      with torch.cuda.stream(s2):
        for k in range(pieces):
          h_out[k * step:(k + 1) * step].copy_(d_out[k * step:(k + 1) * step], non_blocking=True)
  torch.cuda.synchronize()
  dt = (time.perf_counter() - t0) / reps
  return dt * 1e3, n / dt / 1e9


for name, a, b in (('h2d only', True, False), ('d2h only', False, True), ('both', True, True)):
  for pieces in (1, 8):
    run(a, b, reps=2, pieces=pieces)
    ms, gbs = run(a, b, pieces=pieces)
    print(f'{name:9s} pieces={pieces}: {ms:7.2f} ms per 428 MB -> {gbs:6.1f} GB/s per direction')

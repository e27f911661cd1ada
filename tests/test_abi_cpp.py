"""The C ABI exercised WITHOUT PyTorch: tests/cuda/abi_smoke.cu (cudaMalloc buffers -> mlt_gl_attn_fwd / bwd ->
compare with the fp64-oracle fixture) is compiled against include/mlt_attn.h + libmlt_attn.so.  The CPU leg
checks that it compiles and links (every symbol it uses resolves); the GPU leg runs it."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'multimodal-long-transformer-2021_b200')
FIXTURE = os.path.join(ROOT, 'tests', 'golden', 'gl_abi_fixture.bin')


@pytest.fixture(scope='module')
def binary(tmp_path_factory):
  nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
  if not os.path.exists(nvcc) or not os.path.exists(os.path.join(PKG, 'libmlt_attn.so')):
    pytest.skip('nvcc or libmlt_attn.so not available')
  out = str(tmp_path_factory.mktemp('abi') / 'abi_smoke')
  r = subprocess.run([nvcc, '-std=c++17', '-gencode', 'arch=compute_100a,code=sm_100a', '-O1',
                      os.path.join(ROOT, 'tests', 'cuda', 'abi_smoke.cu'), '-I', os.path.join(ROOT, 'include'),
                      '-L', PKG, '-lmlt_attn', '-Xlinker', f'-rpath={PKG}', '-o', out],
                     capture_output=True, text=True, timeout=600)
  assert r.returncode == 0, r.stderr[-3000:]
  return out


def test_cpp_abi_program_compiles_and_links(binary):
  assert os.path.exists(binary) and os.path.getsize(FIXTURE) > 1000


@pytest.mark.gpu
def test_cpp_abi_program_matches_oracle_fixture(binary):
  r = subprocess.run([binary, FIXTURE], capture_output=True, text=True, timeout=300)
  assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
  lines = [l for l in r.stdout.splitlines() if 'max_abs_err' in l]
  assert len(lines) == 24 and all(l.endswith('ok') for l in lines), r.stdout
  assert 'failures 0' in r.stdout

"""The TensorFlow custom-op shim (tf_shim/mlt_ops.cc) cannot be built here (no TensorFlow), but every use of
the C ABI in it is compiled: it is type-checked against include/mlt_attn.h with a minimal stand-in for the TF
API (tests/tf_stub/).  The Python binding (tf_shim/mlt_tf_ops.py) is checked for consistency with the op
registrations: names, gradient registrations, argument counts."""
import ast
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, 'multimodal-long-transformer-2021_b200', 'tf_shim')


def test_shim_type_checks_against_the_c_abi():
  gxx = shutil.which('g++')
  if not gxx:
    pytest.skip('g++ not available')
  r = subprocess.run([gxx, '-std=c++14', '-fsyntax-only', '-Wall', '-I', os.path.join(ROOT, 'tests', 'tf_stub'),
                      '-I', os.path.join(ROOT, 'include'), os.path.join(SHIM, 'mlt_ops.cc')],
                     capture_output=True, text=True, timeout=300)
  assert r.returncode == 0, r.stderr[-3000:]


def _registered_ops():
  src = open(os.path.join(SHIM, 'mlt_ops.cc')).read()
  macros = dict(re.findall(r'#define (MLT_\w+)\s*\\?\n((?:.*\\\n)*.*)\n', src))
  ops = {}
  for m in re.finditer(r'REGISTER_OP\("(\w+)"\)(.*?);', src, re.S):
    body = m.group(2)
    for _ in range(3):
      for k, v in macros.items():
        body = body.replace(k, v)
    ops[m.group(1)] = dict(inputs=re.findall(r'\.Input\("(\w+):', body), outputs=re.findall(r'\.Output\("(\w+):', body),
                           attrs=re.findall(r'\.Attr\("(\w+):', body))
  kernels = set(re.findall(r'MLT_REGISTER\("(\w+)"', src))
  return ops, kernels


def _snake(name):
  return re.sub(r'(?<!^)(?=[A-Z])', '_', name).lower()


def test_python_binding_matches_op_registrations():
  ops, kernels = _registered_ops()
  assert set(ops) == {'MltDenseRelAttn', 'MltDenseRelAttnGrad', 'MltGlAttn', 'MltGlAttnGrad', 'MltGlAttnCompact',
                      'MltGlAttnCompactGrad', 'MltGlSideInputsToCompact'}
  # every attention op has GPU kernels for float and bfloat16; the side-input op is integer-only
  assert kernels == set(ops) - {'MltGlSideInputsToCompact'}
  assert ops['MltGlSideInputsToCompact']['inputs'] == ops['MltGlAttn']['inputs'][10:18]
  assert ops['MltGlSideInputsToCompact']['outputs'][:3] == ops['MltGlAttnCompact']['inputs'][10:13]
  # forward / gradient pairs line up: grad inputs = forward inputs + forward outputs + output gradients
  for fwd, n_dout in (('MltDenseRelAttn', 1), ('MltGlAttn', 2), ('MltGlAttnCompact', 2)):
    f, g = ops[fwd], ops[fwd + 'Grad']
    assert g['inputs'][:len(f['inputs'])] == f['inputs']
    assert g['inputs'][len(f['inputs']):len(f['inputs']) + len(f['outputs'])] == f['outputs']
    assert len(g['inputs']) == len(f['inputs']) + len(f['outputs']) + n_dout
    assert set(f['attrs']) == set(g['attrs'])
  # the reference's dense side inputs, with the names of src/input_utils.py:35-44
  assert ops['MltDenseRelAttn']['inputs'][5:7] == ['att_mask', 'relative_att_ids']
  assert ops['MltGlAttn']['inputs'][10:18] == ['l2l_att_mask', 'l2l_relative_att_ids', 'l2g_att_mask',
                                              'l2g_relative_att_ids', 'g2g_att_mask', 'g2g_relative_att_ids',
                                              'g2l_att_mask', 'g2l_relative_att_ids']
  tree = ast.parse(open(os.path.join(SHIM, 'mlt_tf_ops.py')).read())
  called = {n.func.attr for n in ast.walk(tree) if isinstance(n, ast.Call) and isinstance(n.func, ast.Attribute)
            and isinstance(n.func.value, ast.Name) and n.func.value.id == '_mlt'}
  assert called == {_snake(o) for o in ops}                      # every registered op is bound, nothing else is called
  grads = {d.args[0].value for f in ast.walk(tree) if isinstance(f, ast.FunctionDef) for d in f.decorator_list
           if isinstance(d, ast.Call) and getattr(d.func, 'attr', '') == 'RegisterGradient'}
  assert grads == {'MltDenseRelAttn', 'MltGlAttn', 'MltGlAttnCompact'}
  # INTEGRATION.md only names ops that exist
  doc = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
  for name in re.findall(r'_mlt\.(mlt_\w+)', doc):
    assert name in called, name

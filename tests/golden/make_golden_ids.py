"""Regenerates tests/golden/relative_ids_golden.json.

Runs ONLY in the authoring container (it reads /root/reference, which does not
exist on the GPU box); the JSON it writes is what travels.

The reference's own golden matrices for MmtRelativePositionGenerator live in
``src/feature_utils_test.py`` (smaller case :64-72, larger case :95-108, part ids
:34-35).  TensorFlow cannot be imported in this image, so instead of running the
reference we lift the *expected* literals straight out of its test file with
``ast`` -- they are the reference authors' known answers for this path.
"""
import ast
import json
import pathlib

REF = pathlib.Path('/root/reference/src/feature_utils_test.py')
OUT = pathlib.Path(__file__).with_name('relative_ids_golden.json')


def main():
  tree = ast.parse(REF.read_text())
  cases = []
  for fn in ast.walk(tree):
    if not isinstance(fn, ast.FunctionDef):
      continue
    ctor = None
    expected = None
    seq_len = None
    for node in ast.walk(fn):
      if (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute)
          and node.func.attr == 'MmtRelativePositionGenerator'):
        ctor = {k.arg: ast.literal_eval(k.value) for k in node.keywords}
      if (isinstance(node, ast.Assign) and isinstance(node.targets[0], ast.Name)
          and node.targets[0].id == 'expected'):
        expected = ast.literal_eval(node.value)
      if (isinstance(node, ast.Call) and isinstance(node.func, ast.Attribute)
          and node.func.attr == 'make_relative_att_ids'):
        seq_len = ast.literal_eval(node.args[0])
    if ctor and expected is not None:
      cases.append({'source': f'src/feature_utils_test.py::{fn.name}',
                    'ctor': ctor, 'seq_len': seq_len, 'expected': expected})
  init = {'source': 'src/feature_utils_test.py::test_relative_position_generator_init',
          'ctor': {'num_patch_per_row': 2, 'num_core_layers': 1,
                   'text_relative_pos_max_distance': 3},
          'core_layer_diameter': 3, 'image_part_id': 19, 'text_part_id': 20}
  OUT.write_text(json.dumps({'matrices': cases, 'init': init}, indent=1))
  print(f'wrote {OUT} with {len(cases)} matrices')


if __name__ == '__main__':
  main()

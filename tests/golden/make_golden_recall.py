"""Regenerates tests/golden/recall_golden.json by RUNNING the reference's own
``get_recall_at_k_from_dataframe`` (reference src/prediction_helper.py:30-89).

Runs ONLY in the authoring container (it reads /root/reference).  The reference module imports
TensorFlow at the top, so the function's source is lifted with ``ast`` and executed unmodified with the
three names it needs (``np``, ``pd`` via the dataframe argument, ``collections``): what is pinned is
the reference's own code, not a restatement.
"""
import ast
import collections
import json
import pathlib

import numpy as np
import pandas as pd

REF = pathlib.Path('/root/reference/src/prediction_helper.py')
OUT = pathlib.Path(__file__).with_name('recall_golden.json')


def load_reference_fn():
  tree = ast.parse(REF.read_text())
  fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == 'get_recall_at_k_from_dataframe')
  ns = {'np': np, 'collections': collections}
  exec(compile(ast.Module(body=[fn], type_ignores=[]), str(REF), 'exec'), ns)
  return ns['get_recall_at_k_from_dataframe']


def main():
  ref = load_reference_fn()
  rng = np.random.default_rng(20211018)
  cases = []
  for (n_img, n_txt, drop) in [(7, 7, 0.0), (12, 30, 0.0), (25, 9, 0.0), (10, 14, 0.3), (40, 40, 0.05)]:
    img, txt = np.meshgrid(np.arange(n_img), np.arange(n_txt), indexing='ij')
    img, txt = img.reshape(-1), txt.reshape(-1)
    gt_of_text = rng.integers(0, n_img, n_txt)          # every text has one ground-truth image
    keep = rng.random(img.shape[0]) >= drop               # examples that do not share one pool: missing pairs
    img, txt = img[keep], txt[keep]
    score = rng.random(img.shape[0])
    score[rng.random(img.shape[0]) < 0.1] = 0.5           # ties
    df = pd.DataFrame({'image_index': img, 'text_index': txt, 'gt_image_index': gt_of_text[txt], 'output': score})
    got = ref(df.copy())
    cases.append({'image_index': img.tolist(), 'text_index': txt.tolist(),
                  'gt_image_index': gt_of_text[txt].tolist(), 'output': score.tolist(),
                  'expected': dict(got)})
  OUT.write_text(json.dumps({'source': 'src/prediction_helper.py:30-89 executed in the authoring container',
                             'cases': cases}))
  print(f'wrote {OUT} with {len(cases)} cases')


if __name__ == '__main__':
  main()

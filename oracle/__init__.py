"""CPU oracle for the global-local attention hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is product code: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or the
timed CPU baseline), never as the thing shipped.  The product path
(``multimodal-long-transformer-2021_b200``) never imports this package and
fails loudly when its CUDA library is missing.

Pinning status:

* integer constructors (``feature_oracle``): PINNED by the reference's own
  golden vectors (``src/feature_utils_test.py:25-110`` in the reference tree),
  reproduced in ``tests/golden/`` and checked in ``tests/test_oracle_ids.py``.
* attention outputs / gradients (``attention_oracle``, ``blocked_etc``):
  PARITY UNPINNED.  The arithmetic lives in the un-vendored, un-pinned
  third-party ``etcmodel`` package (google-research monorepo, "clone master",
  reference ``src/README.md:9-10``); it is absent from ``/root/reference`` and
  TensorFlow cannot be imported in this image, so no golden output vector of the
  attention layer exists.  The oracle restates the published ETC algorithm and is
  anchored on the reference call sites
  (``src/modeling/models/mmt_encoder.py:124-135,220-224``).  Mitigation: three
  independent formulations must agree (dense fp64, ETC's blocked algorithm,
  CUDA kernels) plus finite-difference gradient checks.
"""

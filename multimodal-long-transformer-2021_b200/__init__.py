"""B200-native (sm_100a) global-local relative attention.

Drop-in for the attention core the reference reaches through
``etc_layers.RelativeTransformerLayers`` (reference
``src/modeling/models/mmt_encoder.py:124-135,220-224``) and for ETC's
``FusedGlobalLocalAttention`` long-input variant.  The compute lives in
``csrc/`` behind the C ABI declared in ``include/mlt_attn.h``; this package is
the Python host side (torch only for device memory, streams and autograd
plumbing).  There is no CPU fallback: importing ``ops`` without the built
``libmlt_attn.so`` raises.
"""

from . import feature_utils  # noqa: F401
from . import synthetic  # noqa: F401

__all__ = ['feature_utils', 'synthetic', 'sharding']

"""cleantransformer_b200 — B200 (sm_100a) implementation of CleanTransformer's dense
forward/backward + AdamW + DDP gradient all-reduce hot path.

`_lib` is the ctypes view of the C ABI (include/ct_b200.h), `ops` the tensor-level wrappers.
The host-side mirror of the reference's Python classes lives in `transformer`, `optimizer`,
`models.*` and `ddp`.
"""
__version__ = "0.1.0"

"""grandplus_b200 -- B200-native (sm_100a) GRAND+ propagation hot path.

Two parts, both behind the reference's own interfaces and a C ABI (include/grandplus_b200.h):

* ``grandplus_b200.precompute.propagation.Graph(indptr, indices, seed).gfpush_omp(...)``
  -- drop-in for the reference's pybind11 module (/root/reference/precompute/propagation.cpp:8-12).
* ``grandplus_b200.model.random_prop / random_prop_fused / emb`` and the ``Grand_Plus`` mixin
  -- drop-in for /root/reference/model.py:80-87 and /root/reference/model_mag.py:48-55,80-86.

There is no CPU fallback: the CUDA library is required.
"""
__version__ = "0.1.0"

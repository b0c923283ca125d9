"""oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatements of the reference's propagation hot path, used only as the checker:
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  Nothing under ``grand-plus_b200/``
imports it, and the product path raises if its CUDA library is missing.

* :mod:`oracle.gfpush`     -- ctypes binding of ``gfpush_oracle.c`` (GFPush + top-k,
  /root/reference/precompute/graph.h:53-131) and a loader for the reference's own
  pybind11 module when ``oracle/_ref`` has been built (``make -C oracle ref``).
* :mod:`oracle.aggregate`  -- numpy/torch restatement of ``Grand_Plus.random_prop``
  (/root/reference/model.py:80-87) and ``MLP.emb`` (/root/reference/model_mag.py:48-55).
"""

"""ORACLE — TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's HAM hot path (mesh_sfs_optim.py:193-317) used as the parity
checker and as the timed CPU baseline.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this package; nothing under
``fmhr_b200/`` does (tests/test_abi.py::test_product_never_touches_the_oracle enforces it).

Pinning status
--------------
* ``oracle.refmath`` (get_normals / get_matrix / get_radiance / laplacian_smoothing / NCC) restates
  /root/reference/models/utils.py and models/ncc_utils.py.  PINNED: checked against outputs of the
  reference's own functions imported verbatim in the build container; those outputs are committed
  as fixtures under tests/golden/ (generator: oracle/gen_golden.py).
* ``oracle.raster`` (rasterize / interpolate / antialias, C++ in raster_oracle.cpp) restates the
  un-vendored, un-pinned dependency ``nvdiffrast`` (requirements.txt:14).  PARITY UNPINNED: the
  reference has no tests or golden vectors at this boundary; the deterministic rule documented in
  DESIGN.md is the spec.
"""

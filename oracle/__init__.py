"""CPU ORACLE -- test infrastructure, not product code.

A restatement of the reference's per-particle filter arithmetic
(gustavorvillela/mcmh_localization, app/scripts/parallel_utils.py and the glue
arithmetic of app/scripts/amcmh_localizer.py):

* ``oracle/c/mcl_oracle.c``  -- plain C (glibc libm, OpenMP), one function per
  reference function, each citing the reference file:line it follows;
* ``oracle/node_glue.py``    -- NumPy restatement of the node's glue
  (softmax, compute_motion, estimate, map loading, callback order).

Pinned against the UNMODIFIED reference: ``oracle/gen_golden.py`` executes the
reference's numba code in the build container (where /root/reference exists)
and writes ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks the
oracle against those vectors bit-for-bit.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this package.  The
product package ``mcmh_localization_b200`` never imports it and has no CPU
fallback.
"""

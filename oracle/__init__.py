"""CPU oracle for the exact top-k search path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package, and only as the checker or as
the timed CPU baseline.  The product package ``better_search_rag_rust_b200`` never
imports it.

PARITY UNPINNED: the reference holds no golden vectors for this path and cannot be
built here (no cargo/rustc/MPI); see the header of ``bsr_oracle.c``.

``oracle.c_oracle``     ctypes binding of ``bsr_oracle.c`` (the restatement that is timed)
``oracle.numpy_oracle`` independent numpy restatement used to cross-check the C one
"""
from . import c_oracle, numpy_oracle  # noqa: F401

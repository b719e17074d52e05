"""CPU oracle package -- test infrastructure only (see deskew_oracle.py header).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may
import this package.  Parity status: UNPINNED (no reference golden vectors
exist for this path; see deskew_oracle.py).
"""

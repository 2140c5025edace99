"""TEST INFRASTRUCTURE — CPU oracle for the laser-grid point extractor.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package, and only as the checker or the timed
CPU baseline.  The product (cylinder-pose-estimation_b200/) never imports it
and has no CPU path.

Parity status: the reference has no tests, golden vectors or fixtures for this
path -> *parity is pinned only against the reference's own
code executed in the build container* (oracle/import_reference.py, with the
library versions recorded in tests/golden/MANIFEST.json), not against the
pinned wheels of requirements.txt:1-8, and the scikit-image glue is a
restatement (oracle/refshim/skimage).  See DESIGN.md "Oracle".
"""

"""CPU oracle for the brute-force retrieval hot path — TEST INFRASTRUCTURE ONLY.

This package is a CPU restatement (NumPy) of the reference's algorithm for the
path named by BASELINE.json's north_star:

  * metrics.py   - the scalar metric functions of
                   /root/reference/src/geometric_metrics.py:11-129, restated
                   verbatim-in-semantics, plus batched fp32 / fp64 matrices.
  * search.py    - the brute-force scan + stable top-k of
                   /root/reference/src/app_pipeline.py:143-172 and :278-372,
                   and the re-rank/threshold/dedupe of
                   /root/reference/src/image_search.py:115-140,199-219.
  * histogram.py - 8x8x8 colour histograms (RGB / OpenCV-HSV).  The reference
                   has NO histogram code: parity for this
                   function is UNPINNED by the reference and is pinned against
                   OpenCV (cv2.calcHist / cv2.cvtColor) instead.
  * evaluation.py- the evaluation workload of mi_analysis.py (relationship types, per-metric
                   distances, precision / recall threshold counts) on explicit pair lists.
  * synth.py     - seeded synthetic inputs shared by tests and bench.

Pinning: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4 / 8c).  The metric restatement is pinned by executing the
reference's own functions in the build container
(tests/golden/make_golden.py -> tests/golden/metrics_golden.npz, committed) and
by tests/test_oracle.py, which compares every oracle function with those
vectors bit-for-bit.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product package
(image-retrieval-_b200) never imports it and has no CPU fallback.
"""
from . import metrics, search, histogram, synth, evaluation  # noqa: F401

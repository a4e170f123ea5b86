"""CPU oracle for the WEALY retrieval-and-scoring hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker (or as the
timed CPU baseline), never as the thing shipped.  The product package
(``audio-based-lyrics-matching_b200/``, importable as ``wealy_b200``) never imports it and has no
CPU fallback.

Parity status
-------------
* similarity, masked reductions, multi-chunk redux, NT-Xent and CLEWS losses:
  PINNED -- the restatements in ``similarity.py`` / ``masked.py`` / ``losses.py``
  are checked against outputs of the reference's own functions
  (``/root/reference/lib/tensor_ops.py``, ``lib/losses.py``) imported unchanged in
  the build container; the vectors are committed under ``tests/golden/`` together
  with the script that made them (``tests/golden/make_golden.py``).
* per-query ranking, AP / MAP / MR1 and top-k: **parity unpinned** -- the
  reference has not released its evaluator.  The
  restatement in ``evaluator.py`` follows the positives/self definition of
  ``lib/losses.py:40-42``, the distance of ``lib/tensor_ops.py:167-173`` and the
  argument vocabulary of ``lib/audio_dataset/dataset.py:82-86,448-449``; it ships
  in two forms (argsort and rank-count) that are tested equal, plus hand-computed
  known-answer cases.  What the reference does define is pinned: the ranking is
  checked against an argsort of the reference's own "cos" distance goldens, the
  chunk reduction (also with ragged masks) against its distance_tensor_redux goldens.
"""

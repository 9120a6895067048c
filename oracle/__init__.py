"""CPU oracle for the geo-trax extract hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU, the arithmetic that the reference's per-frame loop
(/root/reference/geotrax/extract.py:134-214) delegates to its un-vendored dependencies:

* ``ultralytics>=8.4.80,<9``  (pyproject.toml:56)  -> letterbox / YOLOv8s(-OBB) forward / DFL decode / NMS
* ``stabilo>=1.2.3``          (pyproject.toml:59)  -> gray / resize / mask / ORB / BF-kNN / findHomography / box warp
* OpenCV (``cv2`` 4.13 present in this image) is *called* for ORB / BFMatcher / findHomography, because it is
  the very library the reference's ``stabilo`` calls; torch/torchvision CPU ops are called for conv / nms.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package, and only as the checker or the timed CPU baseline.  The product path
(``geo-trax_b200/``) never imports it and fails loudly when the CUDA library is missing.

Parity pinning status (see DESIGN.md section "Oracle"):
* box-warp semantics, H direction/units/normalisation and the output layout are PINNED against the reference's
  golden files ``data/results-pixel/U_video_cut{,_vid_transf}.txt`` (tests/golden/, tests/test_oracle_golden.py).
* detector and ORB/RANSAC numerics are "parity unpinned" by the reference (it has no tests or vectors for them and
  ships neither the video nor the weights); they are pinned instead against torch / torchvision / OpenCV themselves,
  which are the engines the reference runs.
"""

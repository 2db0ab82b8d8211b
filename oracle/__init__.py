"""CPU oracle for the YOLO-LitePi hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker.  The product
package (``yolo-litepi_b200/``) never imports this package and has no CPU
fallback.

Every function restates a piece of the reference (``/root/reference``) and cites
the file:line it follows.  Pinning status: the reference
ships no tests, golden vectors or KATs, so the restatement is pinned against
*outputs of the reference itself run in the authoring container*
(``tests/golden/make_golden.py`` imports the unmodified ``e2e.py`` and runs the
reference ONNX graph through OpenCV-DNN); the recorded vectors live in
``tests/golden/``.
"""

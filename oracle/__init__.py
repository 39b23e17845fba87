"""CPU oracle for the YOLO-LP post-processing path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  The product (``yolo_lp_b200``) never imports this
package and raises if its CUDA library is missing.

Pinning: the reference ships no tests or golden vectors, and the
greedy-NMS arithmetic lives in the un-vendored dependency ``torchvision.ops.nms``
(``requirements.txt:5`` pins only ``>=0.9.0``).  The oracle is therefore pinned
against outputs of the reference itself, imported from ``/root/reference`` in
the build container with torchvision 0.26.0 (CPU kernel); those outputs are
committed under ``tests/golden/`` together with ``tests/golden/make_golden.py``
that produced them.  ``tests/test_oracle_golden.py`` replays them.

Three pieces: ``lp_oracle`` (numpy fp32 restatement, every function citing its reference lines),
``torch_port`` (torch-CPU port with the real ``torchvision.ops.nms``) and ``stage_ref`` -- the recipe
that stages the UNMODIFIED reference package byte for byte under ``oracle/_ref/`` (git-ignored, shipped
to the GPU box with the snapshot), where it is the timed CPU arm of ``bench.py`` and the checker of the
``-m gpu`` tests that run ``patch.install()`` on the real ``Detect`` / ``non_max_suppression`` /
``Inferer.rescale``.
"""

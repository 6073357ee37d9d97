"""CPU oracle for the OS-CNN + feature-level style-transfer hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it, and there only as the checker (or as the timed CPU baseline),
never as something the CUDA path routes through.

Parity status
-------------
* OS-CNN half (kernel-bank sizing, masks, init replay, conv/BN/ReLU forward + backward,
  shortcut, pooling head): **pinned** against the reference's own modules imported from
  ``/root/reference`` in the build container -- see ``oracle/make_golden.py`` (the generating
  script) and ``tests/golden/*.npz`` (the committed vectors).
* AdaIN / Gram style loss: **parity unpinned** -- the reference has no such operator
  (SURVEY.md F1); the definitions frozen in ``oracle/style.py`` follow SURVEY.md section 8c and are
  pinned only against torch autograd in fp64.
"""

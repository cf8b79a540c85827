"""CPU oracle for the per-sample diagnosis path of Self-Diagnosing GAN.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline``
/ ``--impl reference`` legs may import it, and there only as the checker or as
the timed CPU baseline.  The product package (``self-diagnosing-gan_b200/``)
never imports this package and fails loudly when its CUDA library is missing.

Pinning status (see DESIGN.md "Oracle"):

* ``oracle.scores`` and ``oracle.drs`` restate ``diagan/utils/plot.py:220-249``
  and ``diagan/models/drs.py:7-69`` / ``diagan/trainer/evaluate.py:26-83``.  The
  reference holds no tests or golden vectors for them, so they are pinned
  against OUTPUTS OF THE REFERENCE ITSELF, imported from ``/root/reference`` by
  ``oracle/make_golden.py`` and committed under ``tests/golden/``.
* ``oracle.dcgan`` restates ``diagan/models/mnist.py:155-223`` (eval mode) and is
  pinned the same way (reference module imported through a torch_mimicry shim).
* ``oracle.sngan`` restates torch-mimicry 0.1.16's ``SNGANDiscriminator32/64``
  (pinned in ``requirements.txt:72``; NOT vendored in the reference and not
  installable here).  PARITY UNPINNED: no reference code, test or vector is
  available for it; it follows the published torch-mimicry sources as recalled.
"""

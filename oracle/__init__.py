"""CPU oracle for the empirical-denoiser / thermodynamic-statistics hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import it, and only as the
checker or as the timed CPU baseline -- never as the thing shipped.  The
product path (``physics-of-diffusion-models_b200/``) never imports this package
and fails loudly when its CUDA library is missing.

Parity status: PINNED.  ``oracle/posterior.py`` restates the reference's torch
algorithm (citations are ``file:line`` relative to the reference checkout) and
is checked bit-for-bit / to fp32 round-off against golden vectors produced by
importing the unmodified reference in the build container
(``oracle/make_golden.py`` -> ``tests/golden/*.npz``; ``tests/test_oracle_golden.py``).
The reference holds no tests or golden vectors of its own for this path
(SURVEY.md section 4), so those generated fixtures are the pin.
"""

"""CPU oracle of the sampling hot path - TEST INFRASTRUCTURE ONLY.

A from-scratch restatement (numpy float64 for schedules / integer work, torch-CPU fp32 tensor
algebra for the denoiser and the per-step update) of what the reference computes, each function
citing the reference file:line it follows.  It is pinned two ways:

* ``tests/golden/make_golden.py`` imports the real reference from ``/root/reference`` (with the
  three shims SURVEY section 8(c) lists), runs it on seeded inputs and asserts this oracle agrees;
  the reference's outputs are committed as ``tests/golden/*.npz`` so the pin travels to machines
  where the reference is absent;
* the schedule / respacing / mask hashes are the reference's own, listed in SURVEY section 8(c).

The reference ships no tests or fixtures of its own, so those goldens are the only pin.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  The product (``diffusion-based-motion-style-transfer_b200/``) never
does - it has no CPU path at all.
"""

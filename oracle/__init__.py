"""CPU oracle for the DiFashion conditional denoising step.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product path
(``difashion_b200``) never imports it and fails loudly when its CUDA library is absent.

PARITY UNPINNED: the arithmetic of this path lives in diffusers==0.18.2
(reference ``README.md:27``), which is neither vendored under the reference tree nor
installable here, and the reference ships no tests / golden vectors.  The restatement is
pinned only by builder-made self-checks (parameter count 859 520 964, 686 state-dict
tensors, key names, closed-form scheduler identities) — see ``tests/test_oracle_*.py``.
"""

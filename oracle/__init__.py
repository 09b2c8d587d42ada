"""CPU oracle for the Light-3D-Unet volumetric hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it, and there only as the checker or the
timed CPU arm -- never as the thing shipped.  The product path
(``light-3d-unet-front_b200/light_unet``) raises if its CUDA library is missing.

Each function is a from-scratch restatement (torch CPU fp32 / numpy / plain C)
of the reference algorithm and cites the reference ``file:line`` it follows.

Parity pin: the reference's own tests hold no golden vector for this path
(SURVEY.md section 4), so the oracle is pinned against *outputs of the reference
itself*, imported and run in the build container by
``tests/golden/make_golden.py`` and committed as fixtures under
``tests/golden/``; ``tests/test_oracle_golden.py`` replays them.
"""

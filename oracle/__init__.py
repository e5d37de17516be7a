"""CPU oracle for the VQ codebook quantiser hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product:
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as the
CPU baseline that is timed *beside* the CUDA path -- never as the thing shipped.
The product (``attention-models_b200/vq_b200``) never imports this package and
fails loudly when its CUDA library is missing.

Parity pinning: the reference (pranoyr/attention-models) ships no tests, golden
vectors or fixtures for this path, so the oracle is pinned
against outputs of the reference itself: ``oracle/make_golden.py`` imports the
unmodified ``Codebook`` classes from ``/root/reference`` and writes the fixtures
under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this restatement
against them.
"""

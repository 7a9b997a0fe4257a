"""TEST INFRASTRUCTURE ONLY — CPU oracle for the CFM/DiT hot path.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it, and there only as the checker / reported CPU
baseline — never as the thing measured or shipped.  The product path
(``stabletts_b200``) never imports this package and fails loudly when its CUDA
library is missing.

Parity pinning: the restatement in ``estimator_ref.py`` is checked (a) live
against the reference's own importable ``models.estimator.Decoder`` whenever
``/root/reference`` exists (authoring container), and (b) everywhere against
the committed fixtures under ``tests/golden/`` which were produced by
``oracle/make_golden.py`` from that same unmodified reference.  The ODE
stepping arithmetic belongs to the third-party ``torchdiffeq`` (unpinned in the
reference's requirements.txt:14, absent here): fixed-grid tableaux are restated
from the published algorithm — "parity unpinned" for solver behaviour.
"""

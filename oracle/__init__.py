"""CPU oracle for the Scenario-Wise-Rec hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker.  The product path
(``scenario-wise-rec_b200/``) never imports this package and raises when its CUDA
extension is missing.

Contents
--------
``ref_models``   fp32 torch-CPU restatement of the reference forward passes
                 (EmbeddingLayer, MLP, GateNU and the nine hot-path models), every
                 function citing the reference file:line it follows.  Gradients come
                 from torch autograd over the restated forward.
``ops_ref``      torch-CPU reference of every op of the device program IR (forward
                 AND hand-derived backward), used to check the host-side program
                 builder without a GPU and each CUDA kernel on the GPU.
``make_golden``  imports the unmodified reference from ``/root/reference`` (this
                 container only) and writes ``tests/golden/*.npz``.

Parity pinning: the reference ships no tests/golden vectors, so
the oracle is pinned against outputs of the reference itself, generated here by
``make_golden.py`` and committed under ``tests/golden/``.
"""

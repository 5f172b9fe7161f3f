"""ORACLE -- test infrastructure, NOT product code.

CPU restatement of the reference hot path (maria-korosteleva/Garment-Pattern-Estimation):
  * ``knn_oracle.c`` / ``knn.py``   brute-force kNN with torch_cluster's published fp32 fmaf chain
  * ``thirdparty.py``               the torch_geometric / sparsemax semantics the reference borrows
  * ``model.py``                    nn/net_blocks.py + nn/nets.py (attention, baseline and stitch models) + the 4 active loss
                                    terms, restated
  * ``lstm_decomposed.py``          the LSTM decoder as row GEMMs + cell updates, forward and backward (checked against
                                    torch.nn.LSTM): the specification for replacing the cuDNN call
  * ``ref_stubs.py``                sys.modules stubs that let the UNMODIFIED reference nn/nets.py import
                                    in this container (used only to validate model.py and to generate
                                    tests/golden/* -- /root/reference does not exist on the GPU box)

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs
may import this package, and only as the checker / reported baseline.  The product package
``garment_pattern_estimation_b200`` never imports it and has no CPU fallback.

Parity status: the reference has no tests or golden vectors and its third-party kernels
(torch_geometric, torch_cluster, sparsemax) cannot be installed here, so the third-party semantics are
"parity unpinned" (restated from the published algorithms).  The reference's OWN code (nets.py, net_blocks.py,
metrics/losses.py) IS pinned: ``tests/golden/make_golden.py`` executes it unmodified on top of
``thirdparty.py`` and ``model.py`` must reproduce it bit-for-bit on CPU.
"""

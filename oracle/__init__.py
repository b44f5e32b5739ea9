"""CPU oracle for the v5 Hybrid MAML-STGCN-LSTM hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``weatherforecast_stgcn_maml_b200`` imports
this package.  It may be imported only by ``tests/``, ``__graft_entry__.smoke()``
and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- and there
only as the checker or as the timed CPU baseline, never as the product path.

Parity status: PINNED.  ``oracle/ref_port.py`` is checked (a) against the
unmodified reference files -- ``/root/reference`` in the build container, the
staged archive ``oracle/_ref/reference_py.tar.gz`` (``oracle/build_ref.py``)
on the GPU box -- imported over ``oracle/pyg_shim.py``
(``oracle/make_golden.py``, ``tests/test_oracle_golden.py``), including the
train-mode dropout path with the masks replayed from ATen's CPU generator, and
(b) everywhere against the committed outputs of that run under ``tests/golden/``.  The reference itself
ships no golden vectors or tests, and its one third-party
arithmetic dependency that is absent here, ``torch_geometric.nn.GCNConv``
(version unpinned by the reference's requirements.txt), is restated in
``pyg_shim.py`` from its published algorithm (PyG >= 2.0 defaults).
"""

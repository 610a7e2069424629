"""CPU oracle of the reference's matching head.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this package; nothing under ``fingerprint-matching-code_b200/`` does.

What it is: a plain torch-CPU / numpy / C restatement of ``Net.forward`` of the reference
(``/root/reference/src/model/ngm.py:205-491``) and of every op that forward calls, each function
citing the reference lines it follows.  The reference itself cannot be imported here
(torch_geometric, torch_sparse, torch_spline_conv and pygmtools are absent and there is no
network), so:

* pieces whose reference source still imports (``utils/hungarian.py``, ``utils/feature_align.py``,
  ``src/model/soft_topk.py``, ``src/model/afau.py``, ``src/model/affinity_layer.py``) are PINNED: the
  restatement was checked against them in this container and the agreeing input/output vectors
  are committed under ``tests/golden/`` by ``tests/golden/make_golden.py``;
* pieces owned by absent third-party packages - pygmtools 0.5.3 ``sinkhorn`` (pytorch backend),
  torch_geometric 1.6.3 ``SplineConv`` / ``SAGEConv``, torch_spline_conv 1.2.0 basis/weighting,
  torch_sparse 0.6.8 ``SparseTensor`` mean-spmm - restate the published algorithm.  The reference
  holds no test or golden vector for them: PARITY UNPINNED at those boundaries;
* scipy's ``linear_sum_assignment`` is present in the image, so the LAP restatement
  (``oracle/lap_ref.c`` and ``lap.py``) is pinned against the live scipy on every run.
"""

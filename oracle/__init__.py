"""CPU oracle for the audio->pose hot path.  TEST INFRASTRUCTURE ONLY.

Everything under ``oracle/`` is a CPU restatement (numpy fp64 / torch fp32) of the
reference algorithms on the hot path.  It exists to *check*
the CUDA path.  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it; the product package
``audio-to-motion-generation_b200`` never does and fails loudly without its CUDA library.

Pinning status (see DESIGN.md "Oracle"):
  * mel_oracle, eval_oracle : pinned against the *unmodified* reference functions
    (``pose_video/mel_features.py``, ``motion_evaluation.py``) run in the build container;
    vectors in ``tests/golden/`` were produced by ``oracle/make_golden.py``.
  * model_oracle            : pinned against the unmodified reference classes
    (``model_layers.py``, ``real_motion_model.py``) imported through ``oracle/ref_shim.py``
    with decision D1 applied; the torch_geometric layers are a third-party dependency
    that is absent and unpinned in the reference, so that one boundary is
    **parity unpinned** (restated from PyG's documented GATConv/GraphConv semantics).
"""

"""CPU oracle for the PAAC rollout-and-update hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``paac_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or
as the timed CPU baseline, never as the product path.

It restates, in NumPy (byte/integer work, float64 returns) and torch-CPU
(fp32 like-for-like, fp64 as arbiter), what the reference computes on the path
SURVEY.md section 8(a) lists.  Every function cites the reference file:line it
follows (paths are relative to the upstream checkout).

Pinning status:
  * preprocessing / frame stack / runner protocol: PINNED.  Golden vectors in
    ``tests/golden/`` were produced by ``oracle/make_golden.py`` running the
    reference's own ``environment.FramePool`` / ``ObservationPool`` /
    ``runners.Runners`` / ``emulator_runner.EmulatorRunner`` classes (they
    import unchanged) plus Pillow ``Image.resize(NEAREST)`` (what
    ``scipy.misc.imresize(interp='nearest')`` called).
  * model / loss / clip / RMSProp: the reference's TF-1.0.1 arithmetic lives in
    a dependency that is absent here (TensorFlow 1.0.1, pinned by
    ``pretrained/*/checkpoints/*.meta``).  PINNED TO THE SHIPPED GRAPH:
    ``oracle/tf_graph.py`` interprets the reference's own serialized training
    graph (``-80000000.meta``: forward, loss, TF-autodiff gradient subgraph,
    ``clip_by_global_norm`` subgraph, ``ApplyRMSProp`` nodes) node by node, and
    ``tests/golden/tf_graph_*.npz`` holds its outputs on seeded inputs.  The
    per-op kernels (Conv2D, MatMul, ...) are restated from TF's documented op
    definitions; no TF binary could be run, which is said here plainly.
  * sampling: ``np.random.multinomial`` cannot be driven from injected uniforms;
    the oracle defines inverse-CDF sampling and is checked against multinomial
    only in distribution.  Parity for sampling is "unpinned" beyond that.
"""

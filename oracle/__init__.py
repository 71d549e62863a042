"""oracle/ -- TEST INFRASTRUCTURE ONLY.  Not product code.

CPU/eager restatement of the reference hot path (warp, flow pyramid, STE
quantisation, checkerboard dual prior, Gaussian-conditional and factorised
entropy-bottleneck likelihoods, rate reduction) used as the *checker* for the
CUDA kernels in ``deepvideocodec_b200``.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import anything from here.  The
product package never does, and fails loudly if its CUDA library is missing.

Parity status
-------------
* warp / pyramid / STE / dual prior / rate: restated from files that exist in
  the reference (``dmc/models/layers.py``, ``utils.py``, ``video_model.py``,
  ``dmc/train.py``) and pinned against the reference itself executed in the
  build container -> ``tests/golden/*.npz`` (generator:
  ``tests/golden/make_golden.py``).  **Pinned.**
* Gaussian conditional / entropy bottleneck likelihood arithmetic lives in the
  third-party package ``compressai`` which is absent from ``/root/reference``,
  not installed, and pinned to no version by the reference (no requirements
  file).  ``oracle/compressai`` restates its published algorithm (CompressAI
  >= 1.2 ``entropy_models.py`` / ``ops/bound_ops.py``).  Nothing in the
  reference holds golden vectors for it -> **parity unpinned** for those two
  pieces; compensated by fp64 closed-form checks (scipy) and sum-to-one
  invariants in ``tests/test_oracle_math.py``.
"""

"""Oracle shim of the ``compressai`` surface the reference touches.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Put ``oracle/`` on
``sys.path`` and the stock ``/root/reference/dmc/models/*.py`` import unchanged.
Restated from the published CompressAI algorithm (package absent here; parity
unpinned).  Call sites in the reference: ``dmc/models/video_model.py:7,150,220,
222,232,322,392,394,405`` and ``dmc/models/base_model.py:37,63``.
"""
__version__ = "0.0-oracle-shim"

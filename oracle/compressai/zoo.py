"""Stub of ``compressai.zoo`` (imported by reference ``dmc/train.py:51`` and
``dmc/test.py:22``).  The I-frame codec needs pretrained weights from the
network and is out of scope."""


def cheng2020_anchor(quality, metric="mse", pretrained=False, progress=True, **kwargs):
    raise NotImplementedError(
        "oracle shim: compressai.zoo.cheng2020_anchor is unavailable offline "
        "(I-frame codec is outside the hot path)")

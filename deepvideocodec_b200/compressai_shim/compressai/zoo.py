"""The I-frame codec needs network weights; outside the hot path."""


def cheng2020_anchor(quality, metric="mse", pretrained=False, progress=True, **kwargs):
    raise NotImplementedError(
        "compressai.zoo.cheng2020_anchor is not provided by the deepvideocodec_b200 shim "
        "(I-frame codec, pretrained weights need the network)")

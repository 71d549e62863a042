from deepvideocodec_b200.entropy_models import LowerBound  # noqa: F401

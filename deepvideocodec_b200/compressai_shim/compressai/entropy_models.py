from deepvideocodec_b200.entropy_models import (  # noqa: F401
    EntropyBottleneck, EntropyModel, GaussianConditional)

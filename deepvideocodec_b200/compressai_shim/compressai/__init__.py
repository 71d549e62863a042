"""``compressai`` name shim backed by deepvideocodec_b200 (used only when the
real CompressAI is not installed; see deepvideocodec_b200.patch)."""
__dvc_b200_shim__ = True
__version__ = "0.0-dvc-b200-shim"

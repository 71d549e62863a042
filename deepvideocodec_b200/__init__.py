"""deepvideocodec_b200 -- B200 (sm_100a) hot path of the DMC contextual P-frame
codec: flow warp, checkerboard dual-prior quantisation, Gaussian-conditional /
entropy-bottleneck likelihoods and the rate reduction, as hand-written CUDA
behind the reference's own Python names.

Importing the package never needs a GPU; calling an op needs the built
``libdvc_b200.so`` and CUDA fp32 tensors -- there is no fallback path.
"""
from ._native import (DvcError, LIB_PATH, build_library, built_hash, declared_symbols,  # noqa: F401
                      lib, source_hash)
from .context import (dual_prior_stage_a, dual_prior_stage_b_gc, forward_dual_prior,  # noqa: F401
                      frame_context_compress, frame_context_decompress,
                      frame_context_forward, motion_context_compress,
                      motion_context_decompress, motion_context_forward)
from . import coder  # noqa: F401
from .entropy_models import (EntropyBottleneck, EntropyModel, GaussianConditional,  # noqa: F401
                             LowerBound)
from .graph import GraphedInter  # noqa: F401
from .layers import (bilineardownsacling, flow_pyramid, flow_warp,  # noqa: F401
                     motion_compensation_warps, pack_conv3x3_weight, torch_warp, warp_conv3x3,
                     warp_multi)
from .patch import (install_compressai_shim, motion_compensation_fused, patch,  # noqa: F401
                    unpatch)
from .rate import collect_likelihoods_list, frame_bits, log_sum, rate_finalize  # noqa: F401
from .utils import quantize_around, quantize_ste  # noqa: F401

__version__ = "0.1.0"

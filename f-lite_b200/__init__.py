"""flite_b200 -- B200-native (sm_100a) implementation of the F Lite denoising hot path.

Drop-in for ``f_lite.DiT`` (``/root/reference/f_lite/model.py``) and the sampler loop of
``f_lite.FLitePipeline`` (``/root/reference/f_lite/pipeline.py``); all device work is done by the
hand-written CUDA kernels in ``csrc/`` behind the C ABI declared in ``include/flite_b200.h``.
"""
__version__ = "0.1.0"

from ._lib import FliteError  # noqa: E402,F401
from .model import DiT  # noqa: E402,F401
from .pipeline import APGConfig, FLitePipeline, FLitePipelineOutput, denoise, denoise_step  # noqa: E402,F401

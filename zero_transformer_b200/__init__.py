"""zero_transformer_b200: a B200-native ZeRO-1 GPT pretraining framework.

Layers (top -> bottom): ``train`` / ``generate`` CLIs -> ``Trainer`` ->
``parallel.zero1.Zero1Engine`` (+ ``parallel.fused`` NVLink kernels) ->
``models.gpt`` -> ``ops`` (sm_100a kernels in ``csrc/`` with PyTorch reference
fallbacks on CPU) -> ``data`` / ``utils`` (config, checkpoint, export, metrics).
"""
__version__ = "0.1.0"

from .config import Config, load_config  # noqa: F401
from .models.gpt import GPTConfig, Transformer, model_getter  # noqa: F401

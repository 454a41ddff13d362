"""parler_tts_b200 -- B200-native (sm_100a) implementation of Parler-TTS's generate() hot path.

Public names mirror parler_tts/__init__.py:5-16 of the reference so user code switches with an import:
    from parler_tts_b200 import ParlerTTSForConditionalGeneration, ParlerTTSStreamer, ...
(The directory is spelled with an underscore: `parler-tts_b200` is not an importable Python name.)
"""
__version__ = "0.1.0"

from .configuration import DACConfig, GenerationConfig, ParlerTTSConfig, ParlerTTSDecoderConfig
from .dac_wrapper import DACModel
from .modeling import (
    ParlerTTSCache,
    ParlerTTSForCausalLM,
    ParlerTTSForConditionalGeneration,
    ParlerTTSLogitsProcessor,
    apply_delay_pattern_mask,
    build_delay_pattern_mask,
)
from .incremental import IncrementalDecoder, dac_dependency_radius
from .streamer import ParlerTTSStreamer

__all__ = [
    "ParlerTTSConfig", "ParlerTTSDecoderConfig", "DACConfig", "DACModel", "GenerationConfig", "ParlerTTSForCausalLM",
    "ParlerTTSForConditionalGeneration", "ParlerTTSLogitsProcessor", "apply_delay_pattern_mask",
    "build_delay_pattern_mask", "ParlerTTSStreamer", "IncrementalDecoder", "dac_dependency_radius", "ParlerTTSCache",
]

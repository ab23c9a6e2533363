"""qwen3_asr_mlx_b200: B200-native Qwen3-ASR audio-encoding hot path (log-mel frontend + audio encoder).

Drop-in for that path of gabrimatic/qwen3-asr-mlx: the names exported here are the reference's
(src/qwen3_asr_mlx/__init__.py:16-37) for everything on the path; compute runs in libqasr
(hand-written sm_100a CUDA behind a C ABI, include/qasr.h).
"""

__version__ = "0.1.0"

from .audio import load_audio, log_mel_spectrogram, log_mel_spectrogram_batch, log_mel_spectrogram_packed
from .config import AudioEncoderConfig, ModelConfig, TextDecoderConfig
from .decoder import KVCache, TextDecoder, load_decoder_weights
from .encoder import AudioEncoder, SinusoidalPositionEmbedding, load_encoder_weights
from ._array import DeviceArray
from .generate import prepare_inputs
from .model import LANGUAGE_MAP, Qwen3ASR, TranscriptionResult
from .tokenizer import build_prompt

__all__ = [
    "__version__",
    "load_audio",
    "log_mel_spectrogram",
    "log_mel_spectrogram_batch",
    "log_mel_spectrogram_packed",
    "AudioEncoderConfig",
    "TextDecoderConfig",
    "ModelConfig",
    "TextDecoder",
    "KVCache",
    "load_decoder_weights",
    "AudioEncoder",
    "SinusoidalPositionEmbedding",
    "load_encoder_weights",
    "DeviceArray",
    "Qwen3ASR",
    "TranscriptionResult",
    "LANGUAGE_MAP",
    "prepare_inputs",
    "build_prompt",
]

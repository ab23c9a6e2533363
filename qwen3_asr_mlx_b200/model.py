"""`Qwen3ASR` shell: the reference's public surface (src/qwen3_asr_mlx/model.py:121-275) around the
B200 audio-encoding path.

What runs here is mel + encoder (the hot path) and the long-audio splitting that feeds it.  Text
generation (decoder, sampling, tokenizer — reference decoder.py / generate.py / tokenizer.py) is out
of this path's scope: ``transcribe`` hands the audio embeddings to a pluggable ``decoder_backend``
callable and raises a clear error when none is installed.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass
from pathlib import Path
from typing import Callable, List, Optional, Sequence

import numpy as np

from .audio import SAMPLE_RATE, load_audio
from .config import AudioEncoderConfig
from .encoder import AudioEncoder, load_encoder_weights

# ISO 639-1 hints -> the language names Qwen3-ASR prompts use: the reference's table (model.py:28-96), entry for
# entry (tests/test_reference_pin.py compares it with the reference's own dict).  Hints that are not in the table
# pass through unchanged, as in the reference's _resolve_language (model.py:359-366).
LANGUAGE_MAP = dict(pair.split(":") for pair in (
    "af:Afrikaans ar:Arabic az:Azerbaijani be:Belarusian bg:Bulgarian bn:Bengali bs:Bosnian ca:Catalan cs:Czech cy:Welsh "
    "da:Danish de:German el:Greek en:English es:Spanish et:Estonian fa:Persian fi:Finnish fr:French gl:Galician gu:Gujarati "
    "he:Hebrew hi:Hindi hr:Croatian hu:Hungarian hy:Armenian id:Indonesian is:Icelandic it:Italian ja:Japanese ka:Georgian "
    "kk:Kazakh kn:Kannada ko:Korean lt:Lithuanian lv:Latvian mk:Macedonian ml:Malayalam mn:Mongolian mr:Marathi ms:Malay "
    "my:Burmese ne:Nepali nl:Dutch no:Norwegian pa:Punjabi pl:Polish pt:Portuguese ro:Romanian ru:Russian si:Sinhala "
    "sk:Slovak sl:Slovenian sq:Albanian sr:Serbian sv:Swedish sw:Swahili ta:Tamil te:Telugu th:Thai tl:Filipino tr:Turkish "
    "uk:Ukrainian ur:Urdu uz:Uzbek vi:Vietnamese zh:Chinese").split())


@dataclass
class TranscriptionResult:
    """Same fields as the reference's result (model.py:103-114)."""

    text: str
    language: str
    duration: float


def _find_split_points(samples: np.ndarray, chunk_samples: int, search_samples: int, frame_samples: int = 480) -> List[int]:
    """Sample positions at which to cut long audio (reference model.py:454-513).

    Per-``frame_samples`` RMS energy (float32); for every multiple of ``chunk_samples`` the lowest-
    energy frame within +-``search_samples`` is chosen and the cut snaps to that frame's start.
    Vectorised (the reference evaluates the RMS in a Python list comprehension), same results.
    """
    total = len(samples)
    n_frames = total // frame_samples
    if n_frames == 0:
        return []
    frames = np.asarray(samples[: n_frames * frame_samples]).reshape(n_frames, frame_samples)
    energy = np.sqrt(np.mean(frames ** 2, axis=1)).astype(np.float32)
    radius = search_samples // frame_samples
    points: List[int] = []
    for boundary in range(chunk_samples, total, chunk_samples):
        centre = boundary // frame_samples
        lo, hi = max(0, centre - radius), min(n_frames - 1, centre + radius)
        if lo >= hi:
            points.append(boundary)
        else:
            points.append((int(np.argmin(energy[lo: hi + 1])) + lo) * frame_samples)
    return points


DecoderBackend = Callable[..., str]


class Qwen3ASR:
    """Qwen3-ASR speech model with the audio-encoding path on a B200.

    ``from_pretrained`` / ``transcribe`` / ``warm_up`` / ``close`` / context manager keep the
    reference signatures.  ``encode`` and ``encode_batch`` expose the hot path directly.
    """

    def __init__(self, config: AudioEncoderConfig, encoder: AudioEncoder, decoder_backend: Optional[DecoderBackend] = None,
                 decoder=None):
        self._config = config
        self._encoder = encoder
        self._decoder_backend = decoder_backend
        self._decoder = decoder  # qwen3_asr_mlx_b200.decoder.TextDecoder (prefill on the B200), optional
        self._lock = threading.Lock()

    @classmethod
    def from_pretrained(cls, model_id_or_path, decoder_backend: Optional[DecoderBackend] = None, **kwargs) -> "Qwen3ASR":
        """Load ``config.json`` and the ``audio_tower.*`` weights of ``model.safetensors`` from a local directory or,
        when ``model_id_or_path`` is not one, from the HuggingFace Hub repo of that id (``snapshot_download``, imported
        lazily; extra keyword arguments are forwarded to it) -- reference model.py:151-188.  ``device`` and
        ``load_decoder`` are this implementation's own keywords."""
        from ._hub import model_dir

        device, load_decoder = kwargs.pop("device", None), kwargs.pop("load_decoder", False)
        path = model_dir(model_id_or_path, **kwargs)
        kwargs = {"device": device, "load_decoder": load_decoder}
        config = AudioEncoderConfig.from_pretrained(path)
        encoder = AudioEncoder(config, device=kwargs.get("device"))
        load_encoder_weights(encoder, path)
        decoder = None
        if kwargs.get("load_decoder", False):  # TextDecoder + load_decoder_weights, reference model.py:183-184
            from .config import TextDecoderConfig
            from .decoder import TextDecoder, load_decoder_weights

            decoder = TextDecoder(TextDecoderConfig.from_pretrained(path), device=kwargs.get("device"))
            load_decoder_weights(decoder, path)
        return cls(config, encoder, decoder_backend, decoder)

    # ------------------------------------------------------------------ the hot path
    @staticmethod
    def _as_samples(audio) -> np.ndarray:
        if isinstance(audio, (str, Path)):
            return load_audio(audio)
        samples = np.asarray(audio, dtype=np.float32)
        if samples.ndim != 1:
            raise ValueError(f"Audio array must be 1-D (mono), got shape {samples.shape}")
        return samples

    def encode(self, audio):
        """mel + encoder for one utterance -> ``(1, n_tokens, output_dim)`` (model.py:331-335)."""
        from ._array import DeviceArray

        emb, _ = self._encoder.encode_audio_batch([self._as_samples(audio)])
        return DeviceArray(emb.tensor.unsqueeze(0))

    def encode_batch(self, audios: Sequence):
        """mel + encoder for a batch -> (packed embeddings, token_offsets)."""
        return self._encoder.encode_audio_batch([self._as_samples(a) for a in audios])

    def prefill_batch(self, audios: Sequence, language_tokens: Sequence[int]):
        """Waveforms -> first-token logits and KV cache for a whole batch, every stage on the device:
        mel + encoder (model.py:331-335), ``build_prompt`` (model.py:338-339), ONE ``prepare_inputs`` gather for all
        prompts (generate.py:266) and the decoder prefill (generate.py:269-275).  Needs a ``TextDecoder``.
        ``language_tokens`` are the token ids of ``" " + language name`` (the reference always bakes one in:
        ``Tokenizer.build_prompt(n, language=_resolve_language(...))``, default ``" English"``); the BPE tokenizer is outside
        this path, so the caller encodes the name.  An empty list builds ``language<asr_text>`` with no name, which is NOT
        a prompt the reference ever produces.
        Returns ``(last_logits (B, vocab), KVCache, prompt_offsets (B+1,), audio_token_offsets (B+1,))``."""
        from .generate import prepare_inputs
        from .tokenizer import build_prompt

        if self._decoder is None:
            raise NotImplementedError("no TextDecoder attached: construct Qwen3ASR(..., decoder=TextDecoder) or from_pretrained(load_decoder=True)")
        emb, toffs = self._encoder.encode_audio_batch([self._as_samples(a) for a in audios])
        ids: List[int] = []
        offsets = [0]
        for u in range(len(toffs) - 1):
            ids.extend(build_prompt(int(toffs[u + 1] - toffs[u]), language_tokens))
            offsets.append(len(ids))
        x = prepare_inputs(emb, ids, self._decoder.embed_tokens)  # pads are filled in order: prompt u gets its own rows
        last, cache = self._decoder.prefill(x, offsets)
        return last, cache, np.asarray(offsets, dtype=np.int64), toffs

    def encode_long(self, samples: np.ndarray, chunk_duration: float = 30.0, search_seconds: float = 5.0):
        """Long-audio path of the reference (model.py:382-447) up to the encoder: split at low-energy
        boundaries and encode every segment (per-segment mel max, like model.py:418) as ONE varlen batch.
        Returns ``(packed embeddings, token_offsets, [(start, end) sample span of every segment])``."""
        import torch

        enc = self._encoder
        dev = enc._handle.torch_device
        audio = torch.from_numpy(np.ascontiguousarray(samples, dtype=np.float32)).to(dev)
        cuts = enc.find_split_points(audio, int(chunk_duration * SAMPLE_RATE), int(search_seconds * SAMPLE_RATE))
        spans, prev = [], 0
        for sp in cuts + [len(samples)]:  # the reference's loop, model.py:408-413,441: empty slices are skipped, prev always moves
            if sp > prev:
                spans.append((prev, int(sp)))
            prev = int(sp)
        if any(b - a < 160 for a, b in spans):
            raise ValueError("a segment shorter than one hop (160 samples) cannot be encoded")
        lengths = np.asarray([0] + [b - a for a, b in spans], dtype=np.int64)
        soffs = np.cumsum(lengths)
        contiguous = all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1)) and spans[0][0] == 0
        packed = audio[: spans[-1][1]] if contiguous else torch.cat([audio[a:b] for a, b in spans])
        emb, toffs = enc.encode_packed_audio(packed, soffs)
        return emb, toffs, spans

    # ------------------------------------------------------------------ reference API
    def transcribe(self, audio, language: Optional[str] = None, temperature: float = 0.0, top_p: float = 1.0, top_k: int = 0,
                   repetition_penalty: float = 1.2, max_tokens: Optional[int] = None, repetition_context_size: int = 100,
                   chunk_duration: float = 1200.0) -> TranscriptionResult:
        """Transcribe audio (reference model.py:194-250, 281-447).  The audio-encoding half runs here,
        with all segments of a long file encoded as ONE varlen batch; text generation is delegated."""
        with self._lock:
            samples = self._as_samples(audio)
            if len(samples) == 0:
                return TranscriptionResult(text="", language="Unknown", duration=0.0)
            if self._decoder_backend is None:  # checked before any GPU work is spent on the audio
                raise NotImplementedError(
                    "text generation is outside the B200 audio-encoding path: construct Qwen3ASR with a decoder_backend "
                    "callable(audio_embeddings, n_audio_tokens, language, max_tokens, **sampling) -> str, or use encode()/encode_batch()"
                )
            duration = len(samples) / SAMPLE_RATE
            lang = self._resolve_language(language)
            if duration > chunk_duration:  # strict '>', as in the reference (model.py:313)
                # one upload; the RMS scan + argmin run on the device and the segments are slices of the same buffer
                emb, toffs, spans = self.encode_long(samples, chunk_duration)
                segments = [samples[a:b] for a, b in spans]
            else:
                segments = [samples]
                emb, toffs = self._encoder.encode_audio_batch(segments)
            texts = []
            for i, seg in enumerate(segments):
                seg_tokens = max(256, int(len(seg) / SAMPLE_RATE * 50)) if (max_tokens is None or len(segments) > 1) else max_tokens
                piece = self._decoder_backend(emb[int(toffs[i]): int(toffs[i + 1])], int(toffs[i + 1] - toffs[i]), lang, seg_tokens,
                                              temperature=temperature, top_p=top_p, top_k=top_k, repetition_penalty=repetition_penalty,
                                              repetition_context_size=repetition_context_size)
                if piece:
                    texts.append(piece.strip())
            return TranscriptionResult(text=" ".join(texts), language=lang, duration=duration)

    def _resolve_language(self, language: Optional[str]) -> str:
        if language is None or language.lower() in ("auto", ""):
            return "English"
        return LANGUAGE_MAP.get(language.lower(), language)

    def warm_up(self) -> None:
        """Run 0.5 s of silence through the path once (reference model.py:252-259): builds workspaces and
        lets libqasr capture its launch graph for later calls of the same shape."""
        self._encoder.encode_audio_batch([np.zeros(8000, dtype=np.float32)])

    def close(self) -> None:
        if self._encoder is not None:
            self._encoder.close()
        self._encoder = None
        if getattr(self, "_decoder", None) is not None:
            self._decoder.close()
        self._decoder = None

    def __enter__(self) -> "Qwen3ASR":
        return self

    def __exit__(self, *args) -> None:
        self.close()

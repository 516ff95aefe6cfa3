"""vectorquantizedcpc_b200 -- B200-native (sm_100a) implementation of the VQ-CPC inference hot path:
``Encoder.encode``, ``VQEmbeddingEMA.encode`` and ``Vocoder.generate`` / ``Vocoder.forward`` of
tarepan/VectorQuantizedCPC, behind the reference's own Python method surface.  See DESIGN.md."""
from .model import ConfEncoder, Encoder, VQEmbeddingEMA  # noqa: F401
from .network_vocoder import ConfRNNMSVocoder, ConfVocoder, RNNMSVocoder, Vocoder  # noqa: F401
from . import checkpoint  # noqa: F401  (upstream-format checkpoint ingestion)
from .frontend import ConfPreprocessing, LogMel, wave_to_mel  # noqa: F401  (log-mel front-end, preprocess.py:53-75)

__all__ = ["ConfEncoder", "Encoder", "VQEmbeddingEMA", "ConfVocoder", "ConfRNNMSVocoder", "RNNMSVocoder", "Vocoder",
           "ConfPreprocessing", "LogMel", "wave_to_mel", "checkpoint"]

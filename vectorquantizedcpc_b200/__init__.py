"""vectorquantizedcpc_b200 -- B200-native (sm_100a) implementation of the VQ-CPC inference hot path:
``Encoder.encode``, ``VQEmbeddingEMA.encode`` and ``Vocoder.generate`` / ``Vocoder.forward`` of
tarepan/VectorQuantizedCPC, behind the reference's own Python method surface.  See DESIGN.md."""
from .model import ConfEncoder, Encoder, VQEmbeddingEMA  # noqa: F401
from .network_vocoder import ConfRNNMSVocoder, ConfVocoder, RNNMSVocoder, Vocoder  # noqa: F401
from . import checkpoint  # noqa: F401  (upstream-format checkpoint ingestion)
from .frontend import ConfPreprocessing, LogMel, wave_to_mel  # noqa: F401  (log-mel front-end, preprocess.py:53-75)
from .loudness import integrated_loudness, loudness_normalize  # noqa: F401  (convert.py:57,79-80)
from .pipeline import convert_batch  # noqa: F401  (convert.py:52-83 for a ragged batch)
from .textio import format_txt, save_txt, write_wav  # noqa: F401  (encode.py:48-67 text dump, convert.py:83 wav file)

__all__ = ["ConfEncoder", "Encoder", "VQEmbeddingEMA", "ConfVocoder", "ConfRNNMSVocoder", "RNNMSVocoder", "Vocoder",
           "ConfPreprocessing", "LogMel", "wave_to_mel", "integrated_loudness", "loudness_normalize", "convert_batch", "checkpoint",
           "format_txt", "save_txt", "write_wav"]

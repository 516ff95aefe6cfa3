"""The body of ``convert.py`` (/root/reference/convert.py:52-83) for a ragged batch of utterances, entirely on the GPU:

    ref_loudness = meter.integrated_loudness(wav)                         convert.py:57
    mel          = logmel(preemphasis(wav / max|wav| * 0.999))            convert.py:58-70
    _, _, idx    = encoder.encode(mel)                                    convert.py:76
    output       = vocoder.generate(idx, speaker)                         convert.py:77
    output       = normalize.loudness(output, loudness(output), ref)      convert.py:79-80

The reference runs this loop one utterance at a time with librosa / pyloudnorm on the CPU around two GPU calls; here
every stage takes the whole batch (per-utterance lengths carried through), and nothing returns to the host until the
caller asks for the waveforms."""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import Tensor

from .frontend import LogMel
from .loudness import integrated_loudness, loudness_normalize
from .model import Encoder
from .network_vocoder import Vocoder


def convert_batch(encoder: Encoder, vocoder: Vocoder, waves: Sequence[Tensor], speakers, frontend: Optional[LogMel] = None,
                  match_loudness: bool = True, generator: Optional[torch.Generator] = None,
                  uniforms: Optional[Tensor] = None) -> List[Tensor]:
    """``waves``: 1-D fp32 CUDA tensors (source utterances at ``frontend.conf.sr``); ``speakers``: (B,) target speaker
    ids.  Returns one 1-D waveform per utterance, ``320 * ((1 + N_b // 160 - 2) // 2 + 1)`` samples long."""
    waves = list(waves)
    if not waves:
        return []
    dev = waves[0].device
    fe = frontend if frontend is not None else LogMel().to(dev)
    rate, hop = fe.conf.sr, fe.conf.hop_length
    lens = [int(w.shape[0]) for w in waves]
    batch = torch.zeros(len(waves), max(lens), device=dev)
    for b, w in enumerate(waves):
        if w.dim() != 1 or w.dtype != torch.float32:
            raise ValueError("every wave must be a 1-D float32 tensor")
        batch[b, :lens[b]] = w
    speakers = torch.as_tensor(speakers, dtype=torch.int64, device=dev)
    with torch.no_grad():
        ref = integrated_loudness(batch, rate, lengths=lens) if match_loudness else None          # convert.py:57
        mel = fe(batch, lengths=lens)                                                              # convert.py:58-70
        _, _, idx = encoder.encode(mel)                                                            # convert.py:76
        code_lens = [((1 + n // hop) - 2) // 2 + 1 for n in lens]
        up = 2 * vocoder.conf.rnnms.upsampling_t
        wav = vocoder.generate(idx, speakers, lengths=code_lens, generator=generator, uniforms=uniforms)   # convert.py:77
        out_lens = [up * c for c in code_lens]
        if match_loudness:
            wav, _ = loudness_normalize(wav, ref, rate, lengths=out_lens)                          # convert.py:79-80
    return [wav[b, :out_lens[b]] for b in range(len(waves))]

"""Greedy CTC decoding on the sm_100a kernels.

Replaces (paths relative to the reference repository):
  * speechbrain.decoders.ctc_greedy_decode as called at ssak/infer/general.py:112,
    ssak/infer/speechbrain_infer.py:247 and ssak/train/speechbrain/wav2vec_train.py:70-72
  * torch.argmax(logits, dim=-1) feeding processor.(batch_)decode at ssak/infer/general.py:118 and
    ssak/infer/transformers_infer.py:84-85 (the tokenizer then collapses repeats and drops the
    pad token, site-packages/transformers/models/wav2vec2/tokenization_wav2vec2.py:307-322)
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch

from . import _lib


def greedy_ids(probabilities, n_frames=None, blank_id=0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> (frame_ids int32 [B,T], tokens int32 [B,T] left-packed / -1 padded, lengths int32 [B])."""
    _lib.require_cuda(probabilities, "probabilities")
    if probabilities.dim() != 3:
        raise RuntimeError("probabilities must be [B, T, V]")
    p = probabilities if probabilities.dtype == torch.float32 else probabilities.float()
    if p.stride(2) != 1 and p.size(2) > 1:
        p = p.contiguous()
    B, T, V = p.shape
    dev = p.device
    nf = None
    if n_frames is not None:
        nf = torch.as_tensor(n_frames).to(device=dev, dtype=torch.int32).contiguous()
    ids = torch.empty((B, T), dtype=torch.int32, device=dev)
    out = torch.empty((B, T), dtype=torch.int32, device=dev)
    lens = torch.empty(B, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().ssak_ctc_greedy(p.data_ptr(), B, T, V, p.stride(0), p.stride(1), _lib.ptr(nf),
                                        int(blank_id), ids.data_ptr(), out.data_ptr(), lens.data_ptr(),
                                        torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ssak_ctc_greedy")
    return ids, out, lens


def ctc_greedy_decode(probabilities, seq_lens, blank_id=-1) -> List[List[int]]:
    """SpeechBrain's signature: probabilities [B,T,V], seq_lens relative lengths [B], negative
    blank_id counts from the end of the vocabulary.  -> list of token-id lists."""
    B, T, V = probabilities.shape
    if isinstance(blank_id, int) and blank_id < 0:
        blank_id = V + blank_id
    seq_lens = torch.as_tensor(seq_lens, dtype=torch.float32)
    n = torch.round(seq_lens.to(probabilities.device) * T).to(torch.int32)
    _, out, lens = greedy_ids(probabilities, n, blank_id)
    out, lens = out.cpu(), lens.cpu().tolist()
    return [out[b, : lens[b]].tolist() for b in range(B)]


def argmax_ids(logits) -> torch.Tensor:
    """torch.argmax(logits, dim=-1) for [B,T,V] (or [T,V]) -> int64, first maximal index."""
    squeeze = logits.dim() == 2
    x = logits.unsqueeze(0) if squeeze else logits
    B, T, V = x.shape
    p = _lib.require_cuda(x, "logits")
    p = p if p.dtype == torch.float32 else p.float()
    if p.stride(2) != 1 and V > 1:
        p = p.contiguous()
    ids = torch.empty((B, T), dtype=torch.int32, device=p.device)
    with torch.cuda.device(p.device):
        rc = _lib.lib().ssak_ctc_greedy(p.data_ptr(), B, T, V, p.stride(0), p.stride(1), 0, 0, ids.data_ptr(),
                                        0, 0, torch.cuda.current_stream().cuda_stream)
    _lib.check(rc, "ssak_ctc_greedy")
    ids = ids.to(torch.int64)
    return ids[0] if squeeze else ids


def hf_collapse(logits, pad_token_id: int, n_frames: Optional[torch.Tensor] = None) -> List[List[int]]:
    """What HF's tokenizer keeps of argmax(logits): group-by collapse, pad (= CTC blank) removed."""
    _, out, lens = greedy_ids(logits, n_frames, pad_token_id)
    out, lens = out.cpu(), lens.cpu().tolist()
    return [out[b, : lens[b]].tolist() for b in range(out.size(0))]

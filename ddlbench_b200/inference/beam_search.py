"""Greedy and beam-search decoding for GNMT with length normalisation and coverage penalty
(Wu et al. 2016, eq. 14) — the role of ``pipedream-fork/runtime/translation/seq2seq/inference/
beam_search.py`` (``SequenceGenerator.greedy_search`` / ``beam_search``).

All beams of all sentences advance together as one batch of ``B * beam`` decoder rows; finished
beams are frozen (they only extend with EOS at zero cost), so there is no host-side loop over
hypotheses and a single device->host read at the end.
"""
from __future__ import annotations

import torch

from ..models.gnmt import BOS, EOS, GNMT, PAD


class SequenceGenerator:
    def __init__(self, model: GNMT, beam_size: int = 5, max_seq_len: int = 100,
                 len_norm_factor: float = 0.6, len_norm_const: float = 5.0,
                 cov_penalty_factor: float = 0.1):
        self.model, self.beam, self.max_len = model, beam_size, max_seq_len
        self.lnf, self.lnc, self.cpf = len_norm_factor, len_norm_const, cov_penalty_factor

    @torch.no_grad()
    def greedy_search(self, src, src_len) -> torch.Tensor:
        """-> tokens [B, <=max_len] starting with BOS, EOS-terminated, PAD after EOS."""
        m = self.model
        B = src.shape[1]
        ctx = m.encode(src, src_len)
        tok = torch.full((1, B), BOS, dtype=torch.int64, device=src.device)
        out = [tok[0]]
        done = torch.zeros(B, dtype=torch.bool, device=src.device)
        hidden = None
        for _ in range(self.max_len - 1):
            logits, hidden = m.decode_step(tok, ctx, src_len, hidden)
            nxt = logits.argmax(-1)
            nxt = torch.where(done, torch.full_like(nxt, PAD), nxt)
            out.append(nxt)
            done = done | (nxt == EOS)
            tok = nxt.unsqueeze(0)
            if bool(done.all()):
                break
        return torch.stack(out, dim=1)

    @torch.no_grad()
    def beam_search(self, src, src_len) -> torch.Tensor:
        m, K = self.model, self.beam
        B, dev = src.shape[1], src.device
        ctx = m.encode(src, src_len)                                    # [Ts, B, H]
        Ts = ctx.shape[0]
        ctx = ctx.unsqueeze(2).expand(-1, -1, K, -1).reshape(Ts, B * K, -1)
        slen = src_len.unsqueeze(1).expand(-1, K).reshape(-1)
        scores = torch.zeros(B, K, device=dev)
        scores[:, 1:] = float("-inf")                                   # all beams start identical
        seqs = torch.full((B, K, 1), BOS, dtype=torch.int64, device=dev)
        done = torch.zeros(B, K, dtype=torch.bool, device=dev)
        lengths = torch.ones(B, K, device=dev)
        cov = torch.zeros(B * K, Ts, device=dev)                        # accumulated attention
        hidden = None
        dec = [l for l in m.spec if l[0].startswith("dec.")]
        for _ in range(self.max_len - 1):
            tok = seqs[:, :, -1].reshape(1, B * K)
            # one decoder step that also returns the attention distribution
            h = hidden if hidden is not None else [None] * (len(dec) - 2)
            d = dec[0][1](tok)
            d, attn, h0, probs = dec[1][1](d, ctx, slen, h[0], return_hidden=True)
            new_h = [h0]
            for i, (_, mod, _, _) in enumerate(dec[2:-1], start=1):
                d, hi = mod(d, attn, h[i], return_hidden=True)
                new_h.append(hi)
            logp = torch.log_softmax(dec[-1][1](d)[0].float(), dim=-1).view(B, K, -1)
            V = logp.shape[-1]
            # frozen beams: only EOS continues, at no cost
            frozen = torch.full_like(logp, float("-inf"))
            frozen[..., EOS] = 0.0
            logp = torch.where(done.unsqueeze(-1), frozen, logp)
            cand = (scores.unsqueeze(-1) + logp).view(B, K * V)
            top, idx = cand.topk(K, dim=-1)
            beam_id, tok_id = idx // V, idx % V
            gather = (beam_id + torch.arange(B, device=dev).unsqueeze(1) * K).view(-1)
            seqs = torch.cat([seqs.view(B * K, -1)[gather].view(B, K, -1), tok_id.unsqueeze(-1)], -1)
            was_done = done.view(-1)[gather].view(B, K)
            lengths = lengths.view(-1)[gather].view(B, K) + (~was_done).float()
            done = was_done | (tok_id == EOS)
            scores = top
            cov = cov[gather] + torch.where(was_done.view(-1, 1), torch.zeros_like(cov),
                                            probs[0][gather].float())
            hidden = [tuple(s[:, gather] for s in hs) for hs in new_h]
            if bool(done.all()):
                break
        # final ranking: length normalisation + coverage penalty
        lp = ((self.lnc + lengths) / (self.lnc + 1.0)) ** self.lnf
        mask = (torch.arange(Ts, device=dev).unsqueeze(0) < slen.unsqueeze(1)).float()
        cp = (torch.log(cov.clamp(min=1e-9, max=1.0)) * mask).sum(-1).view(B, K) * self.cpf
        final = scores / lp + cp
        best = final.argmax(-1)
        out = seqs[torch.arange(B, device=dev), best]
        # PAD everything after the first EOS
        eos_pos = (out == EOS).float().cumsum(-1)
        after = (eos_pos - (out == EOS).float()) > 0
        return out.masked_fill(after, PAD)

"""GNMT v2 (Wu et al. 2016 as configured by MLPerf v0.5) — the translation workload the upstream
PipeDream runtime carries next to the image models (``pipedream-fork/runtime/translation/
seq2seq/models/{encoder,decoder,attention,gnmt,seq2seq_base}.py``; not reachable from the
reference's ``run.sh``).

One definition, three uses (same idea as ``models/graph.py`` for the image zoo, generalised to
named multi-tensor edges):

* ``GNMT.forward(src, src_len, tgt)`` — the plain model;
* ``GNMT.split(n)`` — ``n`` pipeline stage modules with tuple interfaces and the tensor names
  that cross each boundary (what the reference ships as hand-generated
  ``translation/models/gnmt/gpus=N/stageK.py`` files);
* ``GNMT.profile_graph()`` — the partitioner's ``graph.txt``.

Layout is time-major ``[T, B, ...]``.  The bidirectional first encoder layer is EMULATED with
two unidirectional LSTMs and a per-sequence time reversal (so it can be cut like any other
layer); the reversal and the padding mask are the two native GNMT kernels of the reference
(``seq2seq/csrc/pack_utils_kernel.cu``, ``pack_utils.cpp``) — here ``ops.revert_varlen`` /
``ops.set_mask`` (``ops/csrc/misc.cu``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops

PAD, UNK, BOS, EOS = 0, 1, 2, 3


# ------------------------------------------------------------------------------------ layers
class RevertVarlen(nn.Module):
    """Time reversal of each sequence within its own length (padding stays at the end)."""

    def forward(self, x, lengths):
        return _Revert.apply(x, lengths)


class _Revert(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, lengths):
        ctx.save_for_backward(lengths)
        return ops.revert_varlen(x, lengths)

    @staticmethod
    def backward(ctx, g):
        (lengths,) = ctx.saved_tensors
        return ops.revert_varlen(g.contiguous(), lengths), None      # the op is an involution


class EmulatedBiLSTM(nn.Module):
    """Bidirectional LSTM as forward LSTM || revert(LSTM(revert(x))).  Padded positions of the
    forward direction see padding input; their outputs are zeroed like ``pad_packed_sequence``."""

    def __init__(self, input_size, hidden_size):
        super().__init__()
        self.fwd = nn.LSTM(input_size, hidden_size)
        self.bwd = nn.LSTM(input_size, hidden_size)
        self.rev = RevertVarlen()

    def forward(self, x, lengths):
        mask = ops.set_mask(lengths, x.shape[0]).to(x.dtype).unsqueeze(-1)        # [T, B, 1]
        a = self.fwd(x)[0] * mask
        b = self.rev(self.bwd(self.rev(x, lengths))[0], lengths)
        return torch.cat([a, b], dim=-1)


class ResidualLSTM(nn.Module):
    def __init__(self, input_size, hidden_size, dropout, residual):
        super().__init__()
        self.drop = nn.Dropout(dropout)
        self.rnn = nn.LSTM(input_size, hidden_size)
        self.residual = residual

    def forward(self, x, hidden=None, extra=None, return_hidden=False):
        h = self.drop(x)
        if extra is not None:
            h = torch.cat([h, extra], dim=-1)
        y, hid = self.rnn(h, hidden)
        if self.residual:
            y = y + x
        return (y, hid) if return_hidden else y


class BahdanauAttention(nn.Module):
    """Normalised additive attention: score = g * v/|v| . tanh(Wq q + Wk k + b)."""

    def __init__(self, query_size, key_size, num_units, normalize=True):
        super().__init__()
        self.linear_q = nn.Linear(query_size, num_units, bias=False)
        self.linear_k = nn.Linear(key_size, num_units, bias=False)
        self.linear_att = nn.Parameter(torch.empty(num_units))
        self.normalize = normalize
        if normalize:
            self.normalize_scalar = nn.Parameter(torch.tensor(1.0 / math.sqrt(num_units)))
            self.normalize_bias = nn.Parameter(torch.zeros(num_units))
        stdv = 1.0 / math.sqrt(num_units)
        nn.init.uniform_(self.linear_att, -0.1, 0.1)
        del stdv

    def forward(self, query, keys, key_len):
        """query [Tq, B, Q], keys [Tk, B, K], key_len [B] -> (context [Tq, B, K], probs [Tq, B, Tk])."""
        q = self.linear_q(query).transpose(0, 1).unsqueeze(2)        # [B, Tq, 1, U]
        k = self.linear_k(keys).transpose(0, 1).unsqueeze(1)         # [B, 1, Tk, U]
        s = q + k
        v = self.linear_att
        if self.normalize:
            s = s + self.normalize_bias
            v = v / v.norm() * self.normalize_scalar
        scores = torch.tanh(s).matmul(v)                              # [B, Tq, Tk]
        mask = ops.set_mask(key_len, keys.shape[0]).t().unsqueeze(1)  # [B, 1, Tk]
        scores = scores.masked_fill(mask == 0, -65504.0)
        probs = F.softmax(scores, dim=-1)
        ctx = torch.bmm(probs, keys.transpose(0, 1))                  # [B, Tq, K]
        return ctx.transpose(0, 1), probs.transpose(0, 1)


class RecurrentAttention(nn.Module):
    """First decoder layer: LSTM whose output queries the encoder context."""

    def __init__(self, input_size, context_size, hidden_size, dropout=0.0):
        super().__init__()
        self.rnn = nn.LSTM(input_size, hidden_size)
        self.attn = BahdanauAttention(hidden_size, context_size, hidden_size, normalize=True)
        self.drop = nn.Dropout(dropout)

    def forward(self, x, context, context_len, hidden=None, return_hidden=False):
        y, hid = self.rnn(x, hidden)
        a, probs = self.attn(y, context, context_len)
        y = self.drop(y)
        return (y, a, hid, probs) if return_hidden else (y, a)


class Classifier(nn.Module):
    def __init__(self, in_features, out_features):
        super().__init__()
        self.classifier = nn.Linear(in_features, out_features)

    def forward(self, x):
        return self.classifier(x)


# ------------------------------------------------------------------------------------- model
class GNMT(nn.Module):
    """``num_layers`` = 4 (``gnmt``) or 8 (``gnmt_large``), hidden 1024, dropout 0.2."""

    def __init__(self, vocab_size: int, hidden_size: int = 1024, num_layers: int = 4,
                 dropout: float = 0.2, share_embedding: bool = False):
        super().__init__()
        H = hidden_size
        self.vocab_size, self.hidden_size, self.num_layers = vocab_size, H, num_layers
        # (name, module, input names, output names) in execution order
        L: List[Tuple[str, nn.Module, Tuple[str, ...], Tuple[str, ...]]] = []
        emb_src = nn.Embedding(vocab_size, H, padding_idx=PAD)
        emb_tgt = emb_src if share_embedding else nn.Embedding(vocab_size, H, padding_idx=PAD)
        L.append(("enc.embed", emb_src, ("src",), ("e",)))
        L.append(("enc.rnn0", EmulatedBiLSTM(H, H), ("e", "src_len"), ("e",)))
        L.append(("enc.rnn1", ResidualLSTM(2 * H, H, dropout, False), ("e",), ("e",)))
        for i in range(2, num_layers):
            L.append((f"enc.rnn{i}", ResidualLSTM(H, H, dropout, True), ("e",), ("e",)))
        L.append(("dec.embed", emb_tgt, ("tgt",), ("d",)))
        L.append(("dec.att_rnn", RecurrentAttention(H, H, H, dropout), ("d", "e", "src_len"),
                  ("d", "attn")))
        L.append(("dec.rnn1", _DecLSTM(2 * H, H, dropout, False), ("d", "attn"), ("d",)))
        for i in range(2, num_layers):
            L.append((f"dec.rnn{i}", _DecLSTM(2 * H, H, dropout, True), ("d", "attn"), ("d",)))
        L.append(("dec.classifier", Classifier(H, vocab_size), ("d",), ("logits",)))
        self.spec = L
        self.layers = nn.ModuleDict({n.replace(".", "_"): m for n, m, _, _ in L})
        for p in self.parameters():
            if p.dim() > 0 and p.numel() > 1:
                nn.init.uniform_(p, -0.1, 0.1)
        with torch.no_grad():
            emb_src.weight[PAD].zero_()
            emb_tgt.weight[PAD].zero_()

    INPUTS = ("src", "src_len", "tgt")

    def forward(self, src, src_len, tgt):
        env = {"src": src, "src_len": src_len, "tgt": tgt}
        _run(self.spec, env)
        return env["logits"]

    # ---- inference helpers (beam search drives the decoder one step at a time)
    def encode(self, src, src_len):
        env = {"src": src, "src_len": src_len}
        _run([l for l in self.spec if l[0].startswith("enc.")], env)
        return env["e"]

    def decode_step(self, tokens, context, src_len, hidden: Optional[list]):
        """tokens [1, B]; hidden: list of per-layer LSTM states (or None) -> (logits [B, V], hidden)."""
        dec = [l for l in self.spec if l[0].startswith("dec.")]
        hidden = hidden if hidden is not None else [None] * (len(dec) - 2)
        new_hidden = []
        d = dec[0][1](tokens)
        d, attn, h, _ = dec[1][1](d, context, src_len, hidden[0], return_hidden=True)
        new_hidden.append(h)
        for i, (_, m, _, _) in enumerate(dec[2:-1], start=1):
            d, h = m(d, attn, hidden[i], return_hidden=True)
            new_hidden.append(h)
        return dec[-1][1](d)[0], new_hidden

    # ---- pipeline stages
    def split(self, num_stages: int, cuts: Optional[Sequence[int]] = None) -> List["GNMTStage"]:
        """Contiguous stages; ``cuts`` are layer indices (exclusive ends), default = balanced by
        parameter count."""
        n = len(self.spec)
        if cuts is None:
            w = [sum(p.numel() for p in m.parameters()) + 1 for _, m, _, _ in self.spec]
            from ..parallel.gpipe import solve_min_max_partition
            sizes = solve_min_max_partition(w, num_stages)
            cuts, acc = [], 0
            for s_ in sizes:
                acc += s_
                cuts.append(acc)
        cuts = list(cuts)
        assert cuts[-1] == n and len(cuts) == num_stages
        stages, start = [], 0
        for k, stop in enumerate(cuts):
            spec = self.spec[start:stop]
            # imports: names read before being (re)defined inside the stage
            inputs = _stage_inputs(spec)
            known = set(inputs)
            for _, _, _, outs in spec:
                known.update(outs)
            # exports: values the rest of the model still reads (pass-through inputs included)
            outputs = ["logits"] if stop == n else \
                sorted(x for x in known if _live_after(self.spec, stop, x))
            stages.append(GNMTStage(spec, inputs, outputs, k))
            start = stop
        # a later stage may read a model input (src_len, tgt) that an earlier stage does not
        # touch: thread it through every stage in between
        for k in range(len(stages) - 1, 0, -1):
            for name in stages[k].input_names:
                prev = stages[k - 1]
                if name not in prev.output_names:
                    prev.output_names.append(name)
                    if name not in prev.input_names and not any(name in o for _, _, _, o in prev.spec):
                        prev.input_names.append(name)
        # the wire format of a boundary is the producer's output order
        for k in range(1, len(stages)):
            prev_out = stages[k - 1].output_names
            extra = [x for x in prev_out if x not in stages[k].input_names]
            assert not [x for x in stages[k].input_names if x not in prev_out]
            stages[k].pass_through = extra            # received but neither read nor forwarded
            stages[k].input_names = list(prev_out)
        return stages

    def profile_graph(self, times_ms: Optional[Dict[str, Tuple[float, float]]] = None):
        """Layer chain with data-dependency edges, as the partitioner's Graph."""
        from ..partition.graph import Graph, Node
        g = Graph()
        last_writer: Dict[str, Node] = {}
        srcs = {}
        for i, name in enumerate(self.INPUTS):
            srcs[name] = Node("node%d" % i, "Input%d" % i, 0.0, 0.0, 0.0, 0.0)
            g.add_node(srcs[name])
            last_writer[name] = srcs[name]
        for j, (name, m, ins, outs) in enumerate(self.spec):
            f, b = (times_ms or {}).get(name, (1.0, 2.0))
            psize = 4.0 * sum(p.numel() for p in m.parameters())
            node = Node("node%d" % (j + len(self.INPUTS)), "%s(%s)" % (type(m).__name__, name), f, b,
                        0.0, psize)
            g.add_node(node)
            for i in ins:
                g.add_edge(last_writer[i], node)
            for o in outs:
                last_writer[o] = node
        return g


class _DecLSTM(ResidualLSTM):
    """Decoder layers 1..: input = concat(dropout(x), attention context)."""

    def forward(self, x, attn, hidden=None, return_hidden=False):
        return super().forward(x, hidden, extra=attn, return_hidden=return_hidden)


def _run(spec, env):
    for _, m, ins, outs in spec:
        y = m(*[env[i] for i in ins])
        if isinstance(m, nn.Embedding) or not isinstance(y, tuple):
            y = (y,)
        for o, v in zip(outs, y):
            env[o] = v


def _stage_inputs(spec) -> List[str]:
    need, have = [], set()
    for _, _, ins, outs in spec:
        for i in ins:
            if i not in have and i not in need:
                need.append(i)
        have.update(outs)
    return need


def _live_after(spec, stop, name) -> bool:
    """Is the value of ``name`` as of position ``stop`` read before being overwritten?"""
    for _, _, ins, outs in spec[stop:]:
        if name in ins:
            return True
        if name in outs:
            return False
    return False


class GNMTStage(nn.Module):
    """One pipeline stage: ``forward(*inputs) -> tuple(outputs)`` in the order of
    ``input_names`` / ``output_names``."""

    def __init__(self, spec, input_names, output_names, index):
        super().__init__()
        self.spec = list(spec)
        self.input_names, self.output_names, self.index = list(input_names), list(output_names), index
        self.layers = nn.ModuleDict({n.replace(".", "_"): m for n, m, _, _ in spec})

    def forward(self, *inputs):
        env = dict(zip(self.input_names, inputs))
        _run(self.spec, env)
        return tuple(env[o] for o in self.output_names)


def gnmt(vocab_size=32320, **kw) -> GNMT:
    return GNMT(vocab_size, num_layers=4, **kw)


def gnmt_large(vocab_size=32320, **kw) -> GNMT:
    return GNMT(vocab_size, num_layers=8, **kw)


class LabelSmoothing(nn.Module):
    """NLL with label smoothing, padding ignored, summed over tokens (reference
    ``seq2seq/train/smoothing.py:8-46``).  On the GPU the fused cross-entropy kernel
    (``ops.xent``) produces loss and d(logits) in one pass."""

    def __init__(self, padding_idx: int = PAD, smoothing: float = 0.1):
        super().__init__()
        self.padding_idx, self.smoothing = padding_idx, smoothing

    def forward(self, logits, target):
        logits = logits.reshape(-1, logits.shape[-1]).float()
        target = target.reshape(-1)
        logp = F.log_softmax(logits, dim=-1)
        nll = -logp.gather(1, target.unsqueeze(1)).squeeze(1)
        smooth = -logp.mean(dim=-1)
        loss = (1.0 - self.smoothing) * nll + self.smoothing * smooth
        return loss.masked_fill(target == self.padding_idx, 0.0).sum()

"""Parallel-text pipeline for GNMT: vocabulary/tokenizer, length-filtered sorted datasets,
bucketing batch sampler, padding collate — the roles of ``pipedream-fork/runtime/translation/
seq2seq/data/{config,tokenizer,dataset,sampler}.py`` — plus a synthetic corpus generator
(there is no WMT16 in a sandbox; the reference downloads it).

Batches are time-major: ``src [Ts, B]``, ``src_len [B]``, ``tgt [Tt, B]``, ``tgt_len [B]``,
padded with ``PAD`` to a multiple of 8 time steps (tensor-core friendly shapes, the same reason
the reference pads its VOCABULARY to a multiple of 8, ``tokenizer.py:33-36``).
"""
from __future__ import annotations

import os
import random
from typing import Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import torch

PAD_TOKEN, UNK_TOKEN, BOS_TOKEN, EOS_TOKEN = "<pad>", "<unk>", "<s>", r"<\s>"
PAD, UNK, BOS, EOS = 0, 1, 2, 3
BPE_JOINER = "@@ "


class Tokenizer:
    def __init__(self, vocab_fname: Optional[str] = None, words: Optional[Iterable[str]] = None,
                 pad_vocab_to: int = 8, separator: str = "@@"):
        vocab = [PAD_TOKEN, UNK_TOKEN, BOS_TOKEN, EOS_TOKEN]
        if vocab_fname is not None:
            with open(vocab_fname, encoding="utf-8") as f:
                vocab += [line.strip().split()[0] for line in f if line.strip()]
        if words is not None:
            vocab += list(words)
        seen, uniq = set(), []
        for w in vocab:
            if w not in seen:
                seen.add(w)
                uniq.append(w)
        i = 0
        while len(uniq) % pad_vocab_to:
            uniq.append("madeupword%04d" % i)
            i += 1
        self.idx2tok = uniq
        self.tok2idx = {w: i for i, w in enumerate(uniq)}
        self.separator = separator

    @property
    def vocab_size(self) -> int:
        return len(self.idx2tok)

    def segment(self, line: str) -> List[int]:
        ids = [self.tok2idx.get(t, UNK) for t in line.strip().split()]
        return [BOS] + ids + [EOS]

    def detokenize(self, ids: Sequence[int], delim: str = " ") -> str:
        toks = []
        for i in ids:
            i = int(i)
            if i == EOS:
                break
            if i in (BOS, PAD):
                continue
            toks.append(self.idx2tok[i])
        text = delim.join(toks)
        return text.replace(self.separator + " ", "").replace(self.separator, "")


class ParallelDataset:
    """Tokenised sentence pairs, filtered to [min_len, max_len] (both sides, incl. BOS/EOS)."""

    def __init__(self, src_fname: str, tgt_fname: str, tokenizer: Tokenizer, min_len: int = 0,
                 max_len: int = 50, max_size: Optional[int] = None, sort: bool = False):
        self.src, self.tgt = [], []
        with open(src_fname, encoding="utf-8") as fs, open(tgt_fname, encoding="utf-8") as ft:
            for ls, lt in zip(fs, ft):
                s, t = tokenizer.segment(ls), tokenizer.segment(lt)
                if min_len <= len(s) <= max_len and min_len <= len(t) <= max_len:
                    self.src.append(s)
                    self.tgt.append(t)
                if max_size and len(self.src) >= max_size:
                    break
        if sort:
            order = sorted(range(len(self.src)), key=lambda i: len(self.src[i]))
            self.src = [self.src[i] for i in order]
            self.tgt = [self.tgt[i] for i in order]
        self.lengths = [max(len(s), len(t)) for s, t in zip(self.src, self.tgt)]

    def __len__(self):
        return len(self.src)

    def __getitem__(self, i):
        return self.src[i], self.tgt[i]


def _pad(seqs: Sequence[Sequence[int]], multiple: int = 8) -> Tuple[torch.Tensor, torch.Tensor]:
    lens = torch.tensor([len(s) for s in seqs], dtype=torch.int64)
    T = int(lens.max())
    T = (T + multiple - 1) // multiple * multiple
    out = torch.full((T, len(seqs)), PAD, dtype=torch.int64)
    for b, s in enumerate(seqs):
        out[:len(s), b] = torch.tensor(s, dtype=torch.int64)
    return out, lens


def collate(pairs, sort_by_src: bool = True):
    """-> (src [Ts,B], src_len, tgt [Tt,B], tgt_len, order) with sequences sorted by decreasing
    source length (``order`` restores the input order)."""
    order = list(range(len(pairs)))
    if sort_by_src:
        order.sort(key=lambda i: -len(pairs[i][0]))
    src, src_len = _pad([pairs[i][0] for i in order])
    tgt, tgt_len = _pad([pairs[i][1] for i in order])
    return src, src_len, tgt, tgt_len, order


class BucketingSampler:
    """Batches of similar length: shuffle, cut into ``bucket_mult`` x batch-size chunks, sort
    each chunk by length, slice batches, shuffle batch order; sharded across ``world`` ranks
    (every rank gets the same number of batches; the tail is dropped)."""

    def __init__(self, dataset: ParallelDataset, batch_size: int, seed: int = 1, world: int = 1,
                 rank: int = 0, bucket_mult: int = 80, shuffle: bool = True):
        self.ds, self.bs, self.seed, self.world, self.rank = dataset, batch_size, seed, world, rank
        self.bucket_mult, self.shuffle, self.epoch = bucket_mult, shuffle, 0
        self.global_bs = batch_size * world
        self.num_batches = len(dataset) // self.global_bs

    def set_epoch(self, epoch: int) -> None:
        self.epoch = epoch

    def __len__(self) -> int:
        return self.num_batches

    def __iter__(self) -> Iterator[List[int]]:
        rng = random.Random(self.seed + self.epoch)
        idx = list(range(len(self.ds)))
        if self.shuffle:
            rng.shuffle(idx)
        idx = idx[:self.num_batches * self.global_bs]
        chunk = self.global_bs * self.bucket_mult
        batches = []
        for c in range(0, len(idx), chunk):
            part = sorted(idx[c:c + chunk], key=lambda i: self.ds.lengths[i])
            for b in range(0, len(part) - self.global_bs + 1, self.global_bs):
                batches.append(part[b:b + self.global_bs])
        if self.shuffle:
            rng.shuffle(batches)
        for gb in batches:
            yield gb[self.rank * self.bs:(self.rank + 1) * self.bs]


class TextLoader:
    def __init__(self, dataset: ParallelDataset, sampler: BucketingSampler, pin: bool = False):
        self.ds, self.sampler, self.pin = dataset, sampler, pin and torch.cuda.is_available()

    def __len__(self):
        return len(self.sampler)

    def __iter__(self):
        for ids in self.sampler:
            batch = collate([self.ds[i] for i in ids])
            if self.pin:
                batch = tuple(t.pin_memory() if isinstance(t, torch.Tensor) else t for t in batch)
            yield batch


# --------------------------------------------------------------------------- synthetic corpus
def generate_synthetic_corpus(out_dir: str, sentences: int = 2000, vocab_words: int = 500,
                              min_len: int = 3, max_len: int = 30, seed: int = 0,
                              name: str = "train") -> Dict[str, str]:
    """A learnable toy translation task: the target sentence is the source reversed, with every
    word mapped through a fixed permutation of the vocabulary.  Writes ``<name>.src``,
    ``<name>.tgt`` and ``vocab.txt``; returns their paths."""
    os.makedirs(out_dir, exist_ok=True)
    rng = random.Random(seed)
    words = ["w%03d" % i for i in range(vocab_words)]
    perm = list(range(vocab_words))
    random.Random(1234).shuffle(perm)
    paths = {"src": os.path.join(out_dir, name + ".src"), "tgt": os.path.join(out_dir, name + ".tgt"),
             "vocab": os.path.join(out_dir, "vocab.txt")}
    with open(paths["src"], "w") as fs, open(paths["tgt"], "w") as ft:
        for _ in range(sentences):
            n = rng.randint(min_len, max_len)
            ids = [rng.randrange(vocab_words) for _ in range(n)]
            fs.write(" ".join(words[i] for i in ids) + "\n")
            ft.write(" ".join(words[perm[i]] for i in reversed(ids)) + "\n")
    if not os.path.exists(paths["vocab"]):
        with open(paths["vocab"], "w") as f:
            f.write("\n".join(words) + "\n")
    return paths

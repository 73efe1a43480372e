"""``translate.py``: translate a tokenized text file with a trained GNMT checkpoint (beam search with
length normalisation and coverage penalty) and optionally score it against a reference — the tool
of ``pipedream-fork/profiler/translation/translate.py:1-227`` with its flags:

    python -m ddlbench_b200.harness.translate --dataset-dir DATA -i newstest.tok.en -o out.tok.de \\
        -m ckpt/ [-r newstest.tok.de] --beam-size 5 --max-seq-len 80 --math fp32

``-m`` is a checkpoint FILE (``{"state_dict": ...}``) or a DIRECTORY of per-stage
``checkpoint.<k>.pth.tar`` files as the pipeline trainer writes them (they are merged, the role of
``runtime/translation/compute_bleu_scores.py``).  The model size is read from the checkpoint.
"""
from __future__ import annotations

import argparse
import os
import time
from typing import List

import torch

from ..data import text
from ..inference.beam_search import SequenceGenerator
from ..models.gnmt import GNMT
from ..utils.bleu import corpus_bleu
from .translation import PLAIN, WMT16


def parse(argv=None):
    p = argparse.ArgumentParser(description="GNMT Translate", formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    d = p.add_argument_group("data setup")
    d.add_argument("--dataset-dir", default="data/wmt16_de_en/", help="directory with the vocabulary")
    d.add_argument("-i", "--input", required=True, help="input file (tokenized)")
    d.add_argument("-o", "--output", required=True, help="output file (tokenized)")
    d.add_argument("-r", "--reference", default=None, help="reference file (for BLEU)")
    d.add_argument("-m", "--model", required=True, help="checkpoint file or directory of stage checkpoints")
    q = p.add_argument_group("inference setup")
    q.add_argument("--batch-size", default=128, type=int)
    q.add_argument("--beam-size", default=5, type=int)
    q.add_argument("--max-seq-len", default=80, type=int)
    q.add_argument("--len-norm-factor", default=0.6, type=float)
    q.add_argument("--cov-penalty-factor", default=0.1, type=float)
    q.add_argument("--len-norm-const", default=5.0, type=float)
    g = p.add_argument_group("general setup")
    g.add_argument("--math", default="fp32", choices=["fp32", "fp16", "bf16"],
                   help="arithmetic type (fp16 of the reference maps to bf16 here)")
    b = g.add_mutually_exclusive_group()
    b.add_argument("--bleu", dest="bleu", action="store_true")
    b.add_argument("--no-bleu", dest="bleu", action="store_false")
    g.set_defaults(bleu=True, cuda=True, cudnn=True, batch_first=False)
    f = g.add_mutually_exclusive_group()
    f.add_argument("--batch-first", dest="batch_first", action="store_true")
    f.add_argument("--seq-first", dest="batch_first", action="store_false")
    c = g.add_mutually_exclusive_group()
    c.add_argument("--cuda", dest="cuda", action="store_true")
    c.add_argument("--no-cuda", dest="cuda", action="store_false")
    n = g.add_mutually_exclusive_group()
    n.add_argument("--cudnn", dest="cudnn", action="store_true")
    n.add_argument("--no-cudnn", dest="cudnn", action="store_false")
    g.add_argument("--print-freq", "-p", default=1, type=int)
    s = p.add_argument_group("distributed setup")
    s.add_argument("--rank", default=0, type=int)
    s.add_argument("--world-size", default=1, type=int)
    s.add_argument("--dist-url", default="tcp://localhost:23456")
    return p.parse_args(argv)


def load_model(path: str, vocab_size: int) -> GNMT:
    """Checkpoint file or directory of stage checkpoints -> GNMT of the stored size."""
    merged = {}
    if os.path.isdir(path):
        k = 0
        while os.path.exists(os.path.join(path, "checkpoint.%d.pth.tar" % k)):
            merged.update(torch.load(os.path.join(path, "checkpoint.%d.pth.tar" % k), map_location="cpu")["state_dict"])
            k += 1
        if not merged:
            raise FileNotFoundError("no checkpoint.<k>.pth.tar under %s" % path)
    else:
        sd = torch.load(path, map_location="cpu")
        merged = sd.get("state_dict", sd)
    emb = [v for k, v in merged.items() if k.endswith("enc_embed.weight")]
    hidden = int(emb[0].shape[1]) if emb else 1024
    # encoder LSTM layers present in the checkpoint: layers.enc_rnn<k>.* (GNMT(num_layers) has num_layers of them)
    enc = {k.split(".")[1] for k in merged if k.startswith("layers.enc_rnn")}
    layers = max(1, len(enc))
    model = GNMT(vocab_size, hidden, layers, 0.0)
    try:
        model.load_state_dict(merged)
    except Exception as e:  # noqa: BLE001
        raise RuntimeError("checkpoint does not match a GNMT(vocab %d, hidden %d, %d layers): %s"
                           % (vocab_size, hidden, layers, str(e)[:300]))
    return model


def main(argv=None) -> float:
    a = parse(argv)
    vocab = os.path.join(a.dataset_dir, WMT16["vocab"])
    if not os.path.exists(vocab):
        vocab = os.path.join(a.dataset_dir, PLAIN["vocab"])
    tok = text.Tokenizer(vocab)
    device = torch.device("cuda" if (a.cuda and torch.cuda.is_available()) else "cpu")
    model = load_model(a.model, tok.vocab_size).to(device).eval()
    if a.math in ("fp16", "bf16") and device.type == "cuda":
        model = model.to(torch.bfloat16)
    gen = SequenceGenerator(model, beam_size=a.beam_size, max_seq_len=a.max_seq_len,
                            len_norm_factor=a.len_norm_factor, len_norm_const=a.len_norm_const,
                            cov_penalty_factor=a.cov_penalty_factor)
    with open(a.input) as f:
        lines = [ln.rstrip("\n") for ln in f]
    ids = [tok.segment(ln) for ln in lines]
    order = sorted(range(len(ids)), key=lambda i: -len(ids[i]))      # long sentences first (static-ish batches)
    hyps: List[str] = [""] * len(ids)
    t0, ntok = time.time(), 0
    for bi, s in enumerate(range(0, len(order), a.batch_size)):
        sel = order[s:s + a.batch_size]
        pairs = [(ids[i], ids[i]) for i in sel]                        # (src, dummy tgt) for the collate helper
        src, sl, _, _, perm = text.collate(pairs)
        out = gen.beam_search(src.to(device), sl.to(device)) if a.beam_size > 1 else \
            gen.greedy_search(src.to(device), sl.to(device))
        for row, j in zip(out.cpu(), perm):
            hyps[sel[j]] = tok.detokenize(row.tolist())
        ntok += int(sl.sum())
        if a.rank == 0 and bi % max(a.print_freq, 1) == 0:
            print("TEST [%d/%d]\tTime %.3f\tDecoder iters %d\tTok/s %.0f" % (
                min(s + a.batch_size, len(order)), len(order), time.time() - t0, a.max_seq_len,
                ntok / max(time.time() - t0, 1e-9)), flush=True)
    os.makedirs(os.path.dirname(os.path.abspath(a.output)) or ".", exist_ok=True)
    with open(a.output, "w") as f:
        f.write("\n".join(hyps) + "\n")
    bleu = float("nan")
    if a.bleu and a.reference:
        with open(a.reference) as f:
            refs = [ln.rstrip("\n") for ln in f][:len(hyps)]
        bleu = corpus_bleu(hyps, refs)[0]
        print("BLEU on test dataset: %.2f" % bleu)
    print("Total time: %.2f s, %d sentences" % (time.time() - t0, len(hyps)))
    return bleu


if __name__ == "__main__":
    main()

"""GNMT training / evaluation harness — the role of ``pipedream-fork/runtime/translation/
main_with_runtime.py`` (pipeline-parallel training with weight stashing, per-stage checkpoints,
its ``Epoch: [e][i/n]`` progress lines) and ``compute_bleu_scores.py`` (merge the per-stage
checkpoints, translate the test set with beam search, report BLEU).

    # single device
    python -m ddlbench_b200.harness.translation --epochs 2 --steps 50
    # 4-stage 1F1B pipeline, one rank per stage
    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 \\
        -m ddlbench_b200.harness.translation --stages 4
    # BLEU of saved checkpoints
    python -m ddlbench_b200.harness.translation --bleu --checkpoint-dir ckpt

Data: ``--data-dir`` with ``train.src/.tgt``, ``val.src/.tgt``, ``vocab.txt``; generated
(synthetic reverse-and-map corpus) when absent — there is no WMT16 in a sandbox.
"""
from __future__ import annotations

import argparse
import os
import time
from typing import Dict

import torch
import torch.distributed as dist
import torch.nn as nn

from ..data import text
from ..inference import SequenceGenerator
from ..models.gnmt import GNMT, PAD, LabelSmoothing
from ..utils.bleu import corpus_bleu


def parse(argv=None):
    ap = argparse.ArgumentParser(description="GNMT (PipeDream translation workload)")
    ap.add_argument("--data-dir", default=None)
    ap.add_argument("--num-layers", type=int, default=4)
    ap.add_argument("--hidden", type=int, default=1024)
    ap.add_argument("--dropout", type=float, default=0.2)
    ap.add_argument("-b", "--batch-size", type=int, default=64)
    ap.add_argument("--max-len", type=int, default=50)
    ap.add_argument("--epochs", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="cap iterations per epoch")
    ap.add_argument("--lr", type=float, default=1e-3)
    ap.add_argument("--optimizer", default="Adam", choices=["Adam", "SGD"])
    ap.add_argument("--smoothing", type=float, default=0.1)
    ap.add_argument("--grad-clip", type=float, default=5.0)
    ap.add_argument("--dtype", default="fp32", choices=["fp32", "bf16"],
                    help="bf16: low-precision model + flat fp32 master weights (Fp16Optimizer)")
    ap.add_argument("--warmup-steps", type=int, default=0, help="GNMT LR schedule: exponential warm-up")
    ap.add_argument("--stages", type=int, default=1)
    ap.add_argument("--beam-size", type=int, default=5)
    ap.add_argument("--eval-sentences", type=int, default=64)
    ap.add_argument("--device", default=None)
    ap.add_argument("--checkpoint-dir", default=None)
    ap.add_argument("--bleu", action="store_true", help="only score checkpoints in --checkpoint-dir")
    ap.add_argument("--print-freq", type=int, default=10)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--synthetic-sentences", type=int, default=4000)
    ap.add_argument("--synthetic-vocab", type=int, default=500)
    return ap.parse_args(argv)


# file names of the reference's WMT16 en-de directory (seq2seq/data/config.py:14-24)
WMT16 = {"vocab": "vocab.bpe.32000",
         "train": ("train.tok.clean.bpe.32000.en", "train.tok.clean.bpe.32000.de"),
         "val": ("newstest_dev.tok.clean.bpe.32000.en", "newstest_dev.tok.clean.bpe.32000.de")}
PLAIN = {"vocab": "vocab.txt", "train": ("train.src", "train.tgt"), "val": ("val.src", "val.tgt")}


def prepare_data(a, rank: int = 0):
    """Tokenizer + train / validation sets from ``--data-dir``: the reference's WMT16 layout
    (``vocab.bpe.32000``, ``train.tok.clean.bpe.32000.{en,de}``, ``newstest_dev...``) or the plain
    ``{train,val}.{src,tgt}`` + ``vocab.txt`` layout; a missing / empty directory gets a synthetic
    corpus (there is no network here).  A directory with other content is never written to."""
    d = a.data_dir or os.path.join(os.environ.get("TMPDIR", "/tmp"), "ddlb_gnmt_data")
    names = WMT16 if os.path.exists(os.path.join(d, WMT16["vocab"])) else PLAIN
    need = not os.path.exists(os.path.join(d, names["train"][0]))
    if dist.is_initialized():
        # rank 0 decides for everybody: another rank may look at the directory while rank 0 is
        # already writing the corpus and would otherwise skip the barrier below
        flag = [need]
        dist.broadcast_object_list(flag, src=0)
        need = flag[0]
    if need:
        # (only rank 0 inspects the directory: the others may see it half written)
        if rank == 0 and os.path.isdir(d) and os.listdir(d) and \
                not os.path.exists(os.path.join(d, PLAIN["vocab"])):
            raise FileNotFoundError("%s has neither the WMT16 files (%s, %s, ...) nor train.src/train.tgt/vocab.txt"
                                    % (d, WMT16["vocab"], WMT16["train"][0]))
        if rank == 0:
            text.generate_synthetic_corpus(d, a.synthetic_sentences, a.synthetic_vocab, 3,
                                           min(30, a.max_len - 2), seed=0, name="train")
            text.generate_synthetic_corpus(d, max(64, a.synthetic_sentences // 20),
                                           a.synthetic_vocab, 3, min(30, a.max_len - 2), seed=1,
                                           name="val")
        if dist.is_initialized():
            dist.barrier()
    tok = text.Tokenizer(os.path.join(d, names["vocab"]))
    train = text.ParallelDataset(os.path.join(d, names["train"][0]), os.path.join(d, names["train"][1]), tok,
                                 max_len=a.max_len)
    val = text.ParallelDataset(os.path.join(d, names["val"][0]), os.path.join(d, names["val"][1]), tok,
                               max_len=a.max_len, sort=True)
    return tok, train, val


def fixed_batches(ds, sampler, T: int):
    """(inputs=(src, src_len, tgt_in), target=tgt_out), every batch padded to T time steps so
    pipeline messages have static shapes."""
    for ids in sampler:
        src, sl, tgt, tl, _ = text.collate([ds[i] for i in ids])

        def fit(x):
            out = torch.full((T, x.shape[1]), PAD, dtype=torch.int64)
            out[:min(T, x.shape[0])] = x[:T]
            return out
        tgt = fit(torch.cat([tgt, torch.full((1, tgt.shape[1]), PAD, dtype=torch.int64)], 0))
        yield (fit(src), sl, tgt[:-1].contiguous()), tgt[1:].contiguous()


def evaluate_bleu(model: GNMT, tok, val, a, device) -> float:
    model.eval()
    gen = SequenceGenerator(model, beam_size=a.beam_size, max_seq_len=a.max_len + 2)
    hyps, refs = [], []
    n = min(len(val), a.eval_sentences)
    for s in range(0, n, a.batch_size):
        pairs = [val[i] for i in range(s, min(n, s + a.batch_size))]
        src, sl, tgt, tl, order = text.collate(pairs)
        out = gen.beam_search(src.to(device), sl.to(device)) if a.beam_size > 1 else \
            gen.greedy_search(src.to(device), sl.to(device))
        for row, i in zip(out.cpu(), order):
            hyps.append(tok.detokenize(row.tolist()))
            refs.append(tok.detokenize(pairs[i][1]))
    return corpus_bleu(hyps, refs)[0]


def save_checkpoint(directory: str, stage: int, module: nn.Module, epoch: int) -> str:
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, "checkpoint.%d.pth.tar" % stage)
    torch.save({"epoch": epoch, "state_dict": module.state_dict()}, path)
    return path


def merge_stage_checkpoints(directory: str, model: GNMT) -> GNMT:
    """Per-stage ``checkpoint.<k>.pth.tar`` files -> one full model (module names are shared
    between ``GNMT.layers`` and ``GNMTStage.layers``, so the state dicts simply union)."""
    merged: Dict[str, torch.Tensor] = {}
    k = 0
    while os.path.exists(os.path.join(directory, "checkpoint.%d.pth.tar" % k)):
        sd = torch.load(os.path.join(directory, "checkpoint.%d.pth.tar" % k), map_location="cpu")
        merged.update(sd["state_dict"])
        k += 1
    if not merged:
        raise FileNotFoundError("no checkpoint.<k>.pth.tar under %s" % directory)
    model.load_state_dict(merged)
    return model


def main(argv=None) -> Dict[str, float]:
    a = parse(argv)
    torch.manual_seed(a.seed)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    use_cuda = torch.cuda.is_available() and a.device != "cpu"
    if use_cuda:
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    device = torch.device("cuda", torch.cuda.current_device()) if use_cuda else torch.device("cpu")
    if world > 1 and not dist.is_initialized():
        dist.init_process_group("nccl" if use_cuda else "gloo")
    tok, train, val = prepare_data(a, rank)
    model = GNMT(tok.vocab_size, a.hidden, a.num_layers, a.dropout)
    crit = LabelSmoothing(PAD, a.smoothing)
    out: Dict[str, float] = {}
    if a.bleu:
        merge_stage_checkpoints(a.checkpoint_dir, model).to(device)
        out["bleu"] = evaluate_bleu(model, tok, val, a, device)
        print("BLEU = %.2f" % out["bleu"])
        return out
    T = (a.max_len + 7) // 8 * 8
    stages = a.stages if world > 1 else 1
    dp = world // stages if world > 1 else 1
    sampler = text.BucketingSampler(train, a.batch_size, seed=a.seed, world=dp,
                                    rank=(rank // stages if stages > 1 else rank) if dp > 1 else 0)
    steps = len(sampler) if a.steps is None else min(a.steps, len(sampler))
    opt_args = {"lr": a.lr}
    if stages == 1:
        model.to(device)
        fp = None
        if a.dtype == "bf16":
            from ..utils.fp_optimizer import Fp16Optimizer
            model.to(torch.bfloat16)
            fp = Fp16Optimizer(model, grad_clip=a.grad_clip or float("inf"))
            opt = getattr(torch.optim, a.optimizer)(fp.fp32_params, **opt_args)
        else:
            opt = getattr(torch.optim, a.optimizer)(model.parameters(), **opt_args)
        sched = None
        if a.warmup_steps:
            from .lr import WarmupMultiStepLR
            sched = WarmupMultiStepLR(opt, iterations=steps * a.epochs, warmup_steps=a.warmup_steps)
        for epoch in range(a.epochs):
            sampler.set_epoch(epoch)
            model.train()
            t0 = time.time()
            ntok = 0
            for i, (inputs, target) in enumerate(fixed_batches(train, sampler, T)):
                if i >= steps:
                    break
                src, sl, tin = (t.to(device, non_blocking=True) for t in inputs)
                target = target.to(device, non_blocking=True)
                loss = crit(model(src, sl, tin), target) / a.batch_size
                if fp is not None and dp == 1:
                    fp.step(loss, opt, sched)         # backward, clip on the fp32 master, update, copy back
                else:
                    opt.zero_grad(set_to_none=True)
                    loss.backward()
                    if dp > 1:
                        for p_ in model.parameters():
                            dist.all_reduce(p_.grad)
                            p_.grad.div_(dp)
                    if a.grad_clip:
                        nn.utils.clip_grad_norm_(model.parameters(), a.grad_clip)
                    opt.step()
                    if sched is not None:
                        sched.step()
                ntok += int((target != PAD).sum()) + int(sl.sum())
                if i % a.print_freq == 0 and rank == 0:
                    print("Epoch: [%d][%d/%d]\tTime %.3f\tLoss %.4f\tTok/s %.0f" % (
                        epoch, i, steps, time.time() - t0, float(loss.detach()), ntok / max(time.time() - t0, 1e-9)))
            out["train_loss"] = float(loss.detach())
            out["tokens_per_sec"] = ntok * dp / max(time.time() - t0, 1e-9)
            if rank == 0:
                print("Epoch %d: %.3f seconds" % (epoch, time.time() - t0))
                out["bleu"] = evaluate_bleu(model, tok, val, a, device)
                print("Epoch %d: BLEU %.2f" % (epoch, out["bleu"]))
                if a.checkpoint_dir:
                    save_checkpoint(a.checkpoint_dir, 0, model, epoch)
        return out
    # ------------------------------------------------------------------ pipeline
    from ..dist.transport import PairGroups
    from ..parallel.pipedream.module_runtime import ModuleStageRuntime, infer_specs
    from ..parallel.pipedream.optimizer import OptimizerWithWeightStashing
    assert world == stages, "one rank per stage (replicate the job for data parallelism)"
    parts = model.split(stages)
    sample, _ = next(fixed_batches(train, text.BucketingSampler(train, a.batch_size, seed=a.seed), T))
    specs = infer_specs(parts, sample, [p.input_names for p in parts], [p.output_names for p in parts])
    groups = PairGroups([(k, k + 1) for k in range(stages - 1)])
    mine = parts[rank].to(device)
    opt = OptimizerWithWeightStashing(a.optimizer, [mine], num_versions=stages - rank, **opt_args)
    rt = ModuleStageRuntime(mine, rank, stages, specs[rank], specs[rank + 1], device, opt,
                            (lambda lg, tg: crit(lg, tg) / a.batch_size), groups=groups)
    for epoch in range(a.epochs):
        sampler.set_epoch(epoch)
        t0 = time.time()

        def on_step(i, loss, epoch=epoch, t0=t0):
            if loss is not None and i % a.print_freq == 0:
                print("Epoch: [%d][%d/%d]\tTime %.3f\tLoss %.4f" % (epoch, i, steps, time.time() - t0,
                                                                   float(loss)))
        rt.run_epoch(fixed_batches(train, sampler, T), steps, on_step)
        if rt.is_last and rt.losses:
            out["train_loss"] = float(rt.losses[-1])
        print("Epoch %d: %.3f seconds" % (epoch, time.time() - t0))
        if a.checkpoint_dir:
            opt.load_new_params()
            save_checkpoint(a.checkpoint_dir, rank, mine, epoch)
    dist.barrier()
    return out


if __name__ == "__main__":
    main()

"""GNMT model, stage splitting, text pipeline, BLEU, beam search (CPU)."""
import os

import pytest
import torch

from ddlbench_b200.data.text import (BucketingSampler, ParallelDataset, TextLoader, Tokenizer,
                                     collate, generate_synthetic_corpus)
from ddlbench_b200.inference import SequenceGenerator
from ddlbench_b200.models.gnmt import BOS, EOS, GNMT, PAD, EmulatedBiLSTM, LabelSmoothing
from ddlbench_b200.utils.bleu import corpus_bleu


def _batch(T=9, B=4, V=64, seed=0):
    g = torch.Generator().manual_seed(seed)
    lens = torch.tensor([9, 7, 4, 2])[:B]
    src = torch.randint(4, V, (T, B), generator=g)
    for b, l in enumerate(lens):
        src[l:, b] = PAD
    tgt = torch.randint(4, V, (T + 1, B), generator=g)
    return src, lens, tgt


def test_emulated_bilstm_matches_packed_bidirectional():
    torch.manual_seed(0)
    H = 8
    em = EmulatedBiLSTM(H, H)
    bi = torch.nn.LSTM(H, H, bidirectional=True)
    with torch.no_grad():
        for a, b in (("weight_ih_l0", em.fwd.weight_ih_l0), ("weight_hh_l0", em.fwd.weight_hh_l0),
                     ("bias_ih_l0", em.fwd.bias_ih_l0), ("bias_hh_l0", em.fwd.bias_hh_l0),
                     ("weight_ih_l0_reverse", em.bwd.weight_ih_l0), ("weight_hh_l0_reverse", em.bwd.weight_hh_l0),
                     ("bias_ih_l0_reverse", em.bwd.bias_ih_l0), ("bias_hh_l0_reverse", em.bwd.bias_hh_l0)):
            getattr(bi, a).copy_(b)
    x = torch.randn(6, 3, H)
    lens = torch.tensor([6, 4, 2])
    packed = torch.nn.utils.rnn.pack_padded_sequence(x, lens)
    ref, _ = torch.nn.utils.rnn.pad_packed_sequence(bi(packed)[0], total_length=6)
    torch.testing.assert_close(em(x, lens), ref, rtol=1e-5, atol=1e-5)


def test_gnmt_forward_backward_and_stage_split_equivalence():
    torch.manual_seed(1)
    m = GNMT(vocab_size=64, hidden_size=16, num_layers=4, dropout=0.0).eval()
    src, lens, tgt = _batch()
    logits = m(src, lens, tgt[:-1])
    assert logits.shape == (9, 4, 64)
    loss = LabelSmoothing(PAD, 0.1)(logits, tgt[1:])
    loss.backward()
    assert all(p.grad is not None for p in m.parameters())
    for n in (2, 3, 4):
        stages = m.split(n)
        assert sum(len(s.spec) for s in stages) == len(m.spec)
        assert set(stages[0].input_names) <= set(GNMT.INPUTS)
        env = {"src": src, "src_len": lens, "tgt": tgt[:-1]}
        vals = [env[k] for k in stages[0].input_names]
        for k, st in enumerate(stages):
            outs = st(*vals)
            if k + 1 < len(stages):
                named = dict(zip(st.output_names, outs))
                vals = [named[i] for i in stages[k + 1].input_names]
        torch.testing.assert_close(outs[0], logits)
    g = m.profile_graph()
    assert len(g.nodes) == len(m.spec) + 3 and "LSTM" in str(g)


def test_text_pipeline(tmp_path):
    p = generate_synthetic_corpus(str(tmp_path), sentences=300, vocab_words=50, max_len=12)
    tok = Tokenizer(p["vocab"])
    assert tok.vocab_size % 8 == 0 and tok.segment("w001 zzz")[:3] == [BOS, tok.tok2idx["w001"], 1]
    assert tok.detokenize([BOS, tok.tok2idx["w001"], tok.tok2idx["w002"], EOS, 5]) == "w001 w002"
    assert Tokenizer(words=["ab@@", "c"]).detokenize([4, 5]) == "abc"
    ds = ParallelDataset(p["src"], p["tgt"], tok, max_len=10)
    assert 0 < len(ds) < 300 and max(ds.lengths) <= 10
    s0 = BucketingSampler(ds, 8, world=2, rank=0)
    s1 = BucketingSampler(ds, 8, world=2, rank=1)
    b0, b1 = list(s0), list(s1)
    assert len(b0) == len(b1) == len(ds) // 16 and not set(sum(b0, [])) & set(sum(b1, []))
    src, sl, tgt, tl, order = next(iter(TextLoader(ds, s0)))
    assert src.shape[0] % 8 == 0 and src.shape[1] == 8 and list(sl) == sorted(sl, reverse=True)
    assert (src[int(sl[0]):, 0] == PAD).all()


def test_bleu():
    ref = ["the cat sat on the mat", "a quick brown fox jumps"]
    assert corpus_bleu(ref, ref)[0] == 100.0
    score, prec, bp = corpus_bleu(["the cat sat on a mat", "a quick brown fox"], ref)
    assert 0 < score < 100 and bp < 1.0 and prec[0] > prec[3]
    assert corpus_bleu(["x y z w"], ["a b c d"])[0] == 0.0


def test_beam_search_and_greedy_learn_copy_task():
    """Train a tiny GNMT on the synthetic reverse-and-map task; decoding must beat chance."""
    torch.manual_seed(0)
    V, T, B = 24, 6, 32
    perm = torch.randperm(V - 4) + 4
    m = GNMT(vocab_size=V, hidden_size=32, num_layers=2, dropout=0.0)
    opt = torch.optim.Adam(m.parameters(), lr=5e-3)
    crit = LabelSmoothing(PAD, 0.0)

    def make(bs):
        src = torch.randint(4, V, (T, bs))
        lens = torch.full((bs,), T)
        body = perm[src.flip(0) - 4]
        tgt = torch.cat([torch.full((1, bs), BOS), body, torch.full((1, bs), EOS)], 0)
        return src, lens, tgt
    for _ in range(600):
        src, lens, tgt = make(B)
        loss = crit(m(src, lens, tgt[:-1]), tgt[1:]) / B
        opt.zero_grad()
        loss.backward()
        opt.step()
    m.eval()
    src, lens, tgt = make(8)
    gen = SequenceGenerator(m, beam_size=3, max_seq_len=T + 2)
    greedy = gen.greedy_search(src, lens)
    beam = gen.beam_search(src, lens)
    want = tgt.t()
    acc_g = (greedy[:, :want.shape[1]] == want[:, :greedy.shape[1]]).float().mean().item()
    acc_b = (beam[:, :want.shape[1]] == want[:, :beam.shape[1]]).float().mean().item()
    assert acc_g > 0.6 and acc_b >= acc_g - 0.1, (acc_g, acc_b)
    assert (beam[:, 0] == BOS).all()


def test_translation_harness_single_process(tmp_path, capsys):
    from ddlbench_b200.harness import translation
    res = translation.main(["--data-dir", str(tmp_path / "data"), "--hidden", "32", "--num-layers", "2",
                            "--batch-size", "16", "--max-len", "16", "--steps", "20", "--epochs", "1",
                            "--lr", "0.01", "--device", "cpu", "--synthetic-sentences", "400",
                            "--synthetic-vocab", "40", "--beam-size", "2", "--eval-sentences", "8",
                            "--checkpoint-dir", str(tmp_path / "ckpt")])
    text_out = capsys.readouterr().out
    assert "Epoch: [0][0/20]" in text_out and "BLEU" in text_out
    assert res["train_loss"] > 0 and "bleu" in res and res["tokens_per_sec"] > 0
    res2 = translation.main(["--data-dir", str(tmp_path / "data"), "--hidden", "32", "--num-layers", "2",
                             "--max-len", "16", "--device", "cpu", "--bleu", "--beam-size", "2",
                             "--eval-sentences", "8", "--batch-size", "16",
                             "--checkpoint-dir", str(tmp_path / "ckpt")])
    assert abs(res2["bleu"] - res["bleu"]) < 1e-6


def test_gnmt_pipeline_1f1b_two_stages(tmp_path):
    import json
    import subprocess
    import sys
    import os
    worker = os.path.join(os.path.dirname(__file__), "gnmt_pipeline_worker.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29577", worker, str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    res = json.load(open(tmp_path / "result.json"))
    assert 0 < res["train_loss"] < 40
    assert os.path.exists(tmp_path / "ckpt" / "checkpoint.0.pth.tar")
    assert os.path.exists(tmp_path / "ckpt" / "checkpoint.1.pth.tar")
    # merged stage checkpoints load into the full model
    from ddlbench_b200.harness import translation
    out = translation.main(["--data-dir", str(tmp_path / "data"), "--hidden", "32", "--num-layers", "2",
                            "--max-len", "16", "--device", "cpu", "--bleu", "--beam-size", "1",
                            "--eval-sentences", "4", "--batch-size", "4",
                            "--checkpoint-dir", str(tmp_path / "ckpt")])
    assert "bleu" in out


def test_translation_harness_bf16_master_weights_and_lr_schedule(tmp_path):
    from ddlbench_b200.harness import translation
    res = translation.main(["--data-dir", str(tmp_path / "data"), "--hidden", "32", "--num-layers", "2",
                            "--batch-size", "16", "--max-len", "16", "--steps", "12", "--epochs", "1",
                            "--lr", "0.01", "--device", "cpu", "--synthetic-sentences", "300",
                            "--synthetic-vocab", "40", "--beam-size", "1", "--eval-sentences", "4",
                            "--dtype", "bf16", "--warmup-steps", "4"])
    assert res["train_loss"] == res["train_loss"] and res["train_loss"] > 0


def test_translation_reads_the_reference_wmt16_layout(tmp_path):
    """--data_dir pointing at the reference's WMT16 directory (vocab.bpe.32000, train.tok.clean.bpe...)
    is read as is; a directory with unrelated content is refused, not filled with synthetic data."""
    import types
    from ddlbench_b200.data import text
    from ddlbench_b200.harness import translation as tr
    src = tmp_path / "plain"
    paths = text.generate_synthetic_corpus(str(src), 60, 40, 3, 10, seed=0, name="train")
    text.generate_synthetic_corpus(str(src), 20, 40, 3, 10, seed=1, name="val")
    wmt = tmp_path / "wmt16_de_en"
    wmt.mkdir()
    (wmt / tr.WMT16["vocab"]).write_text(open(paths["vocab"]).read())
    for split, plain in (("train", ("train.src", "train.tgt")), ("val", ("val.src", "val.tgt"))):
        for dst, name in zip(tr.WMT16[split], plain):
            (wmt / dst).write_text((src / name).read_text())
    a = types.SimpleNamespace(data_dir=str(wmt), max_len=20, synthetic_sentences=10, synthetic_vocab=10)
    tok, train, val = tr.prepare_data(a)
    assert len(train) == 60 and len(val) == 20 and tok.vocab_size % 8 == 0
    assert sorted(os.listdir(wmt)) == sorted([tr.WMT16["vocab"], *tr.WMT16["train"], *tr.WMT16["val"]])
    other = tmp_path / "other"
    other.mkdir()
    (other / "README").write_text("x")
    with pytest.raises(FileNotFoundError):
        tr.prepare_data(types.SimpleNamespace(data_dir=str(other), max_len=20, synthetic_sentences=10, synthetic_vocab=10))


def test_translation_trainer_profiler_and_translate_cli(tmp_path):
    """profiler/translation/{train,translate}.py counterparts: train a tiny GNMT, write the layer
    profile graph, then translate a file with the saved checkpoint (reference flags)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, TMPDIR=str(tmp_path), OMP_NUM_THREADS="2")
    r = subprocess.run([sys.executable, os.path.join(root, "pipedream-fork/profiler/translation/train.py"),
                        "--num-layers", "1", "--hidden-size", "32", "--batch-size", "8", "--steps", "4", "--epochs", "1",
                        "--max-length-train", "16", "--device", "cpu", "--save", "ck", "--profile-directory", "prof",
                        "--eval-sentences", "8", "--beam-size", "2", "--warmup-steps", "2"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "Total forward" in r.stdout and (tmp_path / "prof/gnmt/graph.txt").exists()
    assert (tmp_path / "ck/checkpoint.0.pth.tar").exists()
    data = tmp_path / "ddlb_gnmt_data"
    (tmp_path / "in.txt").write_text("".join(open(data / "val.src").readlines()[:3]))
    (tmp_path / "ref.txt").write_text("".join(open(data / "val.tgt").readlines()[:3]))
    r = subprocess.run([sys.executable, os.path.join(root, "pipedream-fork/profiler/translation/translate.py"),
                        "--dataset-dir", str(data), "-i", "in.txt", "-o", "out/out.txt", "-r", "ref.txt", "-m", "ck",
                        "--beam-size", "2", "--max-seq-len", "20", "--no-cuda", "--math", "fp16"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "BLEU on test dataset" in r.stdout
    assert len((tmp_path / "out/out.txt").read_text().strip().splitlines()) == 3

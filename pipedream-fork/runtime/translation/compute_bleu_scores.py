"""BLEU of a (per-stage) GNMT checkpoint at the reference's path with its flags
(pipedream-fork/runtime/translation/compute_bleu_scores.py: ``--dataset-dir --checkpoint_path
--num_layers --num_stages --beam-size --max-seq-len --module ...``): merges the stage checkpoints,
beam-searches the validation set and prints ``BLEU = x``."""
import argparse
import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "..")))
from ddlbench_b200.harness import translation  # noqa: E402


def main(argv=None):
    ap = argparse.ArgumentParser(description="GNMT BLEU (ddlbench_b200)")
    ap.add_argument("--dataset-dir", default=None, dest="dataset_dir")
    ap.add_argument("--checkpoint_path", required=True)
    ap.add_argument("--num_layers", type=int, default=4)
    ap.add_argument("--num_stages", type=int, default=1)
    ap.add_argument("--batch-size", type=int, default=128, dest="batch_size")
    ap.add_argument("--beam-size", type=int, default=5, dest="beam_size")
    ap.add_argument("--max-seq-len", type=int, default=80, dest="max_seq_len")
    ap.add_argument("--hidden", type=int, default=1024, help="extension: model width (reference: 1024)")
    ap.add_argument("--no-cuda", dest="cuda", action="store_false")
    for flag in ("-i", "--input", "-o", "--output", "-r", "--reference", "--len-norm-factor",
                 "--cov-penalty-factor", "--len-norm-const", "--math", "--print-freq", "-p", "--module"):
        ap.add_argument(flag, default=None)
    for flag in ("--bleu", "--no-bleu", "--batch-first", "--seq-first", "--cudnn", "--no-cudnn"):
        ap.add_argument(flag, action="store_true")
    ap.add_argument("--cuda", dest="cuda", action="store_true", default=True)
    a = ap.parse_args(argv)
    ckpt_dir = a.checkpoint_path if os.path.isdir(a.checkpoint_path) else os.path.dirname(a.checkpoint_path)
    # GNMT layer count in the reference counts encoder layers (decoder has the same number)
    args = ["--bleu", "--checkpoint-dir", ckpt_dir, "--num-layers", str(a.num_layers), "--hidden", str(a.hidden),
            "--beam-size", str(a.beam_size), "--max-len", str(min(a.max_seq_len, 80)), "-b", str(a.batch_size)]
    if a.dataset_dir:
        args += ["--data-dir", a.dataset_dir]
    if not a.cuda:
        args += ["--device", "cpu"]
    return translation.main(args)


if __name__ == "__main__":
    main()

"""The upstream runtime's non-pipelined data-parallel baseline (``pipedream-fork/runtime/
image_classification/main.py``): flags, log-line format (what ``process_output.py`` parses),
checkpoint / resume / forward-only, and a 2-rank Gloo run."""
import os
import re
import subprocess
import sys

from ddlbench_b200.harness import dp_baseline as D

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRAIN_LINE = re.compile(r"^Epoch: \[\d+\]\[\d+/\d+\]\tTime: [\d.]+ \([\d.]+\)\tData: [\d.]+ \([\d.]+\)\t"
                        r"Memory: [\d.]+ \([\d.]+\)\tLoss: [\d.]+ \([\d.]+\)\tPrec@1: [\d.]+ \([\d.]+\)\t"
                        r"Prec@5: [\d.]+ \([\d.]+\)$")


def test_module_names_resolve_like_the_code_generator():
    assert D.resolve_module("models.resnet50.gpus=4", "imagenet") == ("resnet50", "imagenet")
    assert D.resolve_module("models.mnistresnet18.gpus=1", "imagenet") == ("resnet18", "mnist")
    assert D.resolve_module("models.cifar10vgg16.gpus=4", "imagenet") == ("vgg16", "cifar10")
    assert D.resolve_module("vgg11", "cifar10") == ("vgg11", "cifar10")


def test_train_validate_checkpoint_resume_forward_only(tmp_path, capsys):
    common = ["-m", "models.mnistresnet18.gpus=1", "-b", "8", "-s", "-p", "2", "--device", "cpu", "--lr", "0.01",
              "--checkpoint_dir", str(tmp_path)]
    out = D.main(common + ["--epochs", "1", "--num_minibatches", "5"])
    lines = capsys.readouterr().out.splitlines()
    train = [l for l in lines if l.startswith("Epoch: [")]
    assert len(train) == 3 and all(TRAIN_LINE.match(l) for l in train), train
    assert any(l.startswith("Test: [0][0/") for l in lines) and any(l.startswith(" * Prec@1 ") for l in lines)
    assert out["train"]["steps"] == 5 and out["throughput"] > 0
    ck = tmp_path / "checkpoint.pth.tar"
    assert ck.exists()          # (model_best.pth.tar only when the validation accuracy improved)
    # resume: continues at the stored epoch with the stored weights
    out2 = D.main(common + ["--epochs", "2", "--num_minibatches", "3", "--resume", str(ck)])
    lines = capsys.readouterr().out.splitlines()
    assert any("loaded checkpoint" in l and "(epoch 1)" in l for l in lines)
    assert [l for l in lines if l.startswith("Epoch: [")][0].startswith("Epoch: [1][0/3]")
    assert out2["train"]["steps"] == 3
    # forward only: no training lines, no checkpoint written
    os.remove(ck)
    out3 = D.main(common + ["--epochs", "1", "--num_minibatches", "10", "--forward_only"])
    lines = capsys.readouterr().out.splitlines()
    assert not [l for l in lines if l.startswith("Epoch: [")] and "train" not in out3 and not ck.exists()


def test_reference_path_shim_two_ranks_gloo(tmp_path):
    script = os.path.join(ROOT, "pipedream-fork", "runtime", "image_classification", "main.py")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr",
           "127.0.0.1", "--master-port", "29583", script, "-m", "models.mnistresnet18.gpus=2", "-b", "4", "-s",
           "--epochs", "1", "--num_minibatches", "3", "-p", "1", "--device", "cpu", "--lr", "0.01",
           "--checkpoint_dir", str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-3000:]
    assert "2 process(es)" in r.stdout and len([l for l in r.stdout.splitlines() if TRAIN_LINE.match(l)]) == 3
    assert (tmp_path / "checkpoint.pth.tar").exists()

"""run.sh flag parsing / expansion, environment contract, and golden log lines (the formats the
reference's post-processing greps for — SURVEY.md §5.5)."""
import os
import re

import pytest
import torch
import torch.distributed as dist

from ddlbench_b200 import cli
from ddlbench_b200.config import BATCH_SIZES, RunConfig
from ddlbench_b200.harness import logfmt


def _expand(argv):
    return cli.expand(cli.build_parser().parse_args(argv))


def test_run_sh_defaults_and_all_expansion():
    cfgs = _expand([])
    # defaults: mnist / pytorch / 1 GPU / log interval 25 / all six benchmark models / synthetic
    assert [c.arch for c in cfgs] == ["resnet18", "resnet50", "resnet152", "vgg11", "vgg16", "mobilenetv2"]
    assert {c.benchmark for c in cfgs} == {"mnist"} and {c.framework for c in cfgs} == {"pytorch"}
    assert cfgs[0].log_interval == 25 and cfgs[0].synthetic and cfgs[0].gpus == 1
    every = _expand(["-b", "all", "-f", "all", "-m", "all"])
    # 4 frameworks x 4 data sets x 6 models, minus resnet152 under pipedream (disabled in `all`)
    assert len(every) == 4 * 4 * 6 - 4
    assert not [c for c in every if c.framework == "pipedream" and c.arch == "resnet152"]
    assert [c.arch for c in _expand(["-m", "exp2"])] == ["resnet50", "vgg16", "mobilenetv2"]
    # an explicit request for the disabled pair is honoured here
    assert len(_expand(["-f", "pipedream", "-m", "resnet152", "-b", "imagenet"])) == 1
    # -s (real data) has no HIGHRES set
    assert {c.benchmark for c in _expand(["-b", "all", "-s", "-m", "vgg11"])} == {"mnist", "cifar10", "imagenet"}
    # multi-node `all` keeps only the multi-node frameworks
    assert {c.framework for c in _expand(["-f", "all", "-n", "2", "-m", "vgg11"])} == {"horovod", "pipedream"}


def test_flags_reach_the_run_config_and_env_contract():
    (c,) = _expand(["-b", "imagenet", "-f", "pipedream", "-g", "8", "-p", "5", "-m", "resnet50",
                    "--stages", "4x2", "--epochs", "2", "--steps", "7", "--macrobatch"])
    assert (c.gpus, c.log_interval, c.epochs, c.steps_per_epoch) == (8, 5, 2, 7)
    assert c.extra["stages"] == "4x2" and c.extra["macrobatch"] == "1"
    assert c.batch_size == BATCH_SIZES[("imagenet", "pipedream")]["resnet50"] if isinstance(
        BATCH_SIZES.get(("imagenet", "pipedream")), dict) else c.batch_size > 0
    env = c.export_env()
    assert env["LOGINTER"] == "5" and "CORES_GPU" in env and "DATADIR" in env
    assert re.match(r"pipedream - imagenet - resnet50 - batch=\d+", c.header())
    back = cli._cfg_to_argv(c, cli.build_parser().parse_args(["-f", "pipedream"]))
    assert "--stages" in back and "4x2" in back and "--macrobatch" in back


def test_golden_log_lines():
    assert logfmt.creating_model("resnet50") == "=> creating model 'resnet50'"
    assert re.fullmatch(r"train \| 1/3 epoch \(40%\) \| 1234\.568 samples/sec \(estimated\) \| "
                        r"mem \(GB\): \d+\.\d{3} \(\d+\.\d{3}\) / \d+\.\d{3}",
                        logfmt.train_progress(1, 3, 40.0, 1234.5678))
    assert logfmt.valid_progress(97.0, 10.0) == "valid | 97% | 10.000 samples/sec (estimated)"
    assert logfmt.epoch_summary(2, 3, 0.5, 100.0, 0.25, 0.9) == \
        "2/3 epoch | train loss:0.500 100.000 samples/sec | valid loss:0.250 accuracy:0.900"
    assert logfmt.final_summary(0.91234, 100.0, 12.5) == \
        "valid accuracy: 0.9123 | 100.000 samples/sec, 12.500 sec/epoch (average)"
    assert logfmt.test_set(0.1234, 98.5) == "\nTEST SET: Average loss: 0.1234, Accuracy: 98.50%\n"
    line = logfmt.pd_train(0, 5, 60, 0.1, 0.2, 0.001, 1.5, 2.5, 6.9, 7.0, 1.0, 2.0, 3.0, 4.0)
    assert line == ("Epoch: [0][5/60]\tTime: 0.100 (0.200)\tEpoch time [hr]: 0.001 (0.001)\t"
                    "Memory: 1.500 (2.500)\tLoss: 6.9000 (7.0000)\tPrec@1: 1.000 (2.000)\t"
                    "Prec@5: 3.000 (4.000)")
    assert logfmt.pd_train_nonlast(0, 5, 60, 1.5, 2.5) == "Epoch: [0][5/60]\tMemory: 1.500 (2.500)"
    assert logfmt.pd_test(0, 1, 4, 0.1, 0.1, 0.0, 0.0, 1.0, 1.0, 5.0, 5.0, 9.0, 9.0).startswith("Test: [0][1/4]\tTime: 0.100")
    assert logfmt.pd_prec(12.5, 50.0) == " * Prec@1 12.500 Prec@5 50.000"
    assert logfmt.pd_epoch_time(3, 12.3456) == "Epoch 3: 12.346 seconds"
    assert logfmt.pd_lr(1, 100, 0.01) == "Epoch: 1 Step 100 \tLearning rate: 0.010000"
    # the post-processing regexes of the reference find what they look for
    assert re.search(r"Epoch (\d+): ([\d.]+) seconds", logfmt.pd_epoch_time(3, 12.3456))
    assert re.search(r"\* Prec@1 ([\d.]+) Prec@5 ([\d.]+)", logfmt.pd_prec(12.5, 50.0))


def test_hvd_allreduce_equals_serial_sum_single_process():
    """hvd API on a 1-rank gloo group: average / sum semantics and the optimizer wrapper."""
    from ddlbench_b200.parallel import dp
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29631", rank=0, world_size=1)
    try:
        x = torch.arange(6.0)
        assert torch.equal(dp.allreduce(x.clone(), name="x"), x)
        assert dp.size() == 1 and dp.rank() == 0 and dp.local_rank() == 0
        m = torch.nn.Linear(3, 2)
        opt = dp.DistributedOptimizer(torch.optim.SGD(m.parameters(), lr=0.1),
                                      named_parameters=m.named_parameters())
        dp.broadcast_parameters(m.state_dict(), root_rank=0)
        dp.broadcast_optimizer_state(opt, root_rank=0)
        before = m.weight.detach().clone()
        m(torch.ones(1, 3)).sum().backward()
        opt.step()
        assert not torch.equal(before, m.weight)
    finally:
        dist.destroy_process_group()


def test_run_sh_pipedream_plans_and_runs_on_cpu(tmp_path):
    """run.sh front end -> profile -> partition -> generate -> 2-rank PipeDream run (Gloo)."""
    import os
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = ["bash", os.path.join(root, "run", "run", "run.sh"), "-b", "mnist", "-f", "pipedream", "-g", "2",
           "-m", "resnet18", "--steps", "4", "--epochs", "1", "--eval-steps", "1", "--device", "cpu",
           "--workdir", str(tmp_path / "plan"), "--batch-size", "8", "--out", str(tmp_path / "out"),
           "--master-port", "29733"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    text = r.stdout
    assert re.search(r"Epoch 0: [\d.]+ seconds", text) and " * Prec@1 " in text
    assert "device-timed |" in text
    plan = tmp_path / "plan"
    produced = [p.name for p in plan.rglob("*") if p.is_file()]
    assert any(n == "graph.txt" for n in produced) and any(n.endswith("_conf.json") for n in produced)
    outs = list((tmp_path / "out").glob("*/slurm.out"))
    assert outs and "pipedream - mnist - resnet18" in outs[0].read_text()


def test_script_level_hyperparameter_flags():
    (c,) = _expand(["-b", "imagenet", "-f", "pipedream", "-g", "4", "-m", "vgg16", "--lr", "0.02",
                    "--momentum", "0.8", "--wd", "1e-5", "--lr_policy", "polynomial", "--lr_warmup", "2",
                    "--forward_only", "--recompute", "--fp16", "--loss_scale", "128"])
    assert (c.lr, c.momentum, c.weight_decay) == (0.02, 0.8, 1e-5)
    assert c.extra["lr_policy"] == "polynomial" and c.extra["lr_warmup"] == "2" and c.extra["forward_only"] == "1"
    back = cli._cfg_to_argv(c, cli.build_parser().parse_args(["-f", "pipedream"]))
    again = cli.expand(cli.build_parser().parse_args(back))[0]
    assert (again.lr, again.momentum, again.weight_decay) == (0.02, 0.8, 1e-5)
    assert again.extra["lr_policy"] == "polynomial" and again.extra["forward_only"] == "1"
    # defaults still come from the reference's per-script table when the flags are absent
    (d,) = _expand(["-b", "mnist", "-f", "pytorch", "-m", "resnet18"])
    assert (d.lr, d.momentum, d.weight_decay) == (0.01, 0.5, 0.0)


def test_reference_named_script_entry_points(tmp_path, monkeypatch):
    """benchmark/<ds>/<ds>_<framework>.py keep the reference scripts' command line (-a, -s, --lr ...)
    and environment contract."""
    import os
    import subprocess
    import sys
    from ddlbench_b200 import scripts
    monkeypatch.setenv("BATCH_SIZE", "16")
    monkeypatch.setenv("EPOCHS", "2")
    a = scripts.parse("imagenet", ["-a", "vgg16", "-s", "-1", "--lr", "0.05"])
    cfg = scripts.config_from_args("imagenet", "horovod", a)
    assert (cfg.benchmark, cfg.arch, cfg.lr, cfg.batch_size, cfg.epochs) == ("highres", "vgg16", 0.05, 16, 2)
    assert not scripts.config_from_args("cifar10", "pytorch", scripts.parse("cifar10", ["-s", "1"])).synthetic
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for ds in ("mnist", "cifar10", "imagenet"):
        for fw in ("pytorch", "horovod", "gpipe"):
            assert os.path.isfile(os.path.join(root, "benchmark", ds, "%s_%s.py" % (ds, fw)))
    env = dict(os.environ, BATCH_SIZE="8", EPOCHS="1", STEPS="2", EVAL_STEPS="1", LOGINTER="1")
    r = subprocess.run([sys.executable, os.path.join(root, "benchmark", "mnist", "mnist_pytorch.py"),
                        "-a", "resnet18", "-s", "0", "--device", "cpu"], env=env, capture_output=True,
                       text=True, timeout=600, cwd=str(tmp_path))
    assert r.returncode == 0, r.stderr[-1500:]
    assert "=> creating model 'resnet18'" in r.stdout and "valid accuracy:" in r.stdout


def test_pipedream_fork_path_shims(tmp_path):
    """The reference's tool paths (pipedream-fork/{profiler,optimizer,runtime}/...) exist here with
    its flags: profile -> partition -> generate -> run two gloo ranks from the generated conf."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fork = os.path.join(root, "pipedream-fork")
    env = dict(os.environ, STEPS="2", EVAL_STEPS="1")

    def run(*cmd):
        r = subprocess.run([sys.executable, *cmd], cwd=tmp_path, env=env, capture_output=True, text=True,
                           timeout=600)
        assert r.returncode == 0, r.stdout + r.stderr
        return r.stdout

    run(os.path.join(fork, "profiler/image_classification/main.py"), "-a", "cifar10_resnet18", "-b", "4",
        "--profile_directory", "pp/", "-s", "1")
    out = run(os.path.join(fork, "optimizer/optimizer_graph_hierarchical.py"), "-f",
              "pp/cifar10_resnet18/graph.txt", "-n", "2", "-b", "1e10", "-s", "16e9",
              "--use_memory_constraint", "-o", "pp/part")
    assert "Time taken by single-stage pipeline" in out
    run(os.path.join(fork, "optimizer/convert_graph_to_model.py"), "-f", "pp/part/gpus=2.txt", "-n", "resnet18",
        "-a", "resnet18", "-o", "pp/models/resnet18/gpus=2", "--stage_to_num_ranks_map", "0:1,1:1")
    conf = tmp_path / "pp/models/resnet18/gpus=2/mp_conf.json"
    assert conf.exists()
    port = str(29700 + os.getpid() % 200)
    procs = [subprocess.Popen([sys.executable, os.path.join(fork, "runtime/image_classification/main_with_runtime.py"),
                               "--module", "models.cifar10_resnet18.gpus=2", "-b", "8", "--config_path", str(conf),
                               "--rank", str(r), "--local_rank", str(r), "--master_addr", "127.0.0.1",
                               "--master_port", port, "--distributed_backend", "gloo", "--epochs", "1"],
                              cwd=tmp_path, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert "Loss:" in outs[1] and "Prec@1" in outs[1]


def test_translation_path_shims(tmp_path):
    """pipedream-fork/runtime/translation/{main_with_runtime,compute_bleu_scores}.py with the
    reference's flags: train a tiny GNMT for a few minibatches, write the stage checkpoint, score it."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    d = os.path.join(root, "pipedream-fork", "runtime", "translation")
    env = dict(os.environ, TMPDIR=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(d, "main_with_runtime.py"), "--module", "models.gnmt.gpus=1", "-b", "16",
                        "--epochs", "1", "--num_minibatches", "4", "--hidden", "32", "--max-length-train", "24",
                        "--checkpoint_dir", "ck", "--distributed_backend", "gloo", "-s", "--lr", "0.001"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "Epoch 0: BLEU" in r.stdout, r.stdout + r.stderr
    assert (tmp_path / "ck" / "checkpoint.0.pth.tar").exists()
    r = subprocess.run([sys.executable, os.path.join(d, "compute_bleu_scores.py"), "--checkpoint_path", "ck", "--num_layers", "4",
                        "--num_stages", "1", "--hidden", "32", "--max-seq-len", "24", "--no-cuda", "--module",
                        "models.gnmt.gpus=1", "-i", "in", "-o", "out"],
                       cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "BLEU = " in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("rel", ["optimizer/optimizer.py", "optimizer/inference_optimizer_graph.py", "runtime/driver.py",
                                 "runtime/launch.py", "profiler/translation/train.py"])
def test_reference_path_wrappers_start(rel):
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "pipedream-fork", rel), "--help"], capture_output=True,
                       text=True, timeout=300, cwd="/tmp")
    assert r.returncode == 0 and "usage:" in r.stdout, r.stdout + r.stderr

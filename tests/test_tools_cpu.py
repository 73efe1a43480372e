"""Auxiliary tools: tracer, inference / flat partitioners, log post-processing, driver YAML,
activation logger, sparsity analysis, layer summary."""
import os
import pickle

import pytest
import torch
import torch.nn as nn

from ddlbench_b200 import driver
from ddlbench_b200.models import build_model
from ddlbench_b200.partition import flat_optimizer, inference_optimizer
from ddlbench_b200.partition.graph import Graph, Node
from ddlbench_b200.partition.tracer import to_layer_graph, trace_to_graph
from ddlbench_b200.utils import process_output, sparsity, summary
from ddlbench_b200.utils.activation_logger import ActivationAndGradientLogger


class Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.c1 = nn.Conv2d(3, 8, 3, padding=1)
        self.r = nn.ReLU()
        self.c2 = nn.Conv2d(8, 8, 3, padding=1)
        self.pool = nn.AdaptiveAvgPool2d(1)
        self.fc = nn.Linear(8, 4)

    def forward(self, x):
        y = self.r(self.c1(x))
        y = self.c2(y) + y
        return self.fc(torch.flatten(self.pool(y), 1))


def test_fx_tracer_builds_graph_and_layer_graph():
    m = Block().eval()
    g = trace_to_graph(m)
    descs = [n.node_desc for n in g.nodes.values()]
    assert descs[0] == "Input0" and any(d.startswith("Conv2d(3, 8") for d in descs)
    assert "Add()" in descs and any(d.startswith("View") for d in descs)
    add = next(n for n in g.nodes.values() if n.node_desc == "Add()")
    assert len(g.in_edges[add.node_id]) == 2
    lg = to_layer_graph(m, (3, 8, 8), 4).eval()
    x = torch.randn(2, 3, 8, 8)
    with torch.no_grad():
        assert torch.allclose(lg(x), m(x), atol=1e-6)


def _chain(costs):
    g = Graph()
    prev = None
    for i, c in enumerate(costs):
        n = Node("node%d" % i, "L%d()" % i, c * 1000, 0.0, 1e3, 1e3)
        g.add_node(n)
        if prev is not None:
            g.add_edge(prev, n)
        prev = n
    return g


def test_inference_optimizer_meets_throughput_constraint():
    g = _chain([1.0, 1.0, 1.0, 1.0])
    t, lat, splits, reps = inference_optimizer.main(4, None, 1.5, 1e12, graph=g, out=lambda *a: None)
    assert t is not None and t <= 1.5 and lat >= 4.0 - 1e-9
    t2, lat2, _, _ = inference_optimizer.main(1, None, 1.5, 1e12, graph=_chain([1.0] * 4),
                                              out=lambda *a: None)
    assert t2 is None          # one worker cannot sustain an input every 1.5 s


def test_flat_optimizer_on_csv_profile(tmp_path):
    p = tmp_path / "prof.csv"
    p.write_text("layer,fwd,bwd,act,param\n" + "\n".join("L%d,%f,%f,1000,1000" % (i, c, c)
                                                          for i, c in enumerate([1, 2, 3, 4, 5, 6])))
    rows = flat_optimizer.read_profile_csv(str(p))
    assert len(rows) == 6
    t, ends, reps = flat_optimizer.optimize_chain(rows, 3, 1e12, straight_pipeline=True)
    assert ends[-1] == 6 and reps == [1, 1, 1]
    assert t == pytest.approx(0.018, rel=1e-6)      # 2,4,6 | 8,10 | 12 ms -> max 18 ms
    g = flat_optimizer.convert_profile_to_graph(rows)
    assert len(g.nodes) == 6 and len(flat_optimizer.compress_graph_branches(g).nodes) == 1


def test_process_output_table():
    log = ["Epoch 0: 36.0 seconds", " * Prec@1 10.000 Prec@5 50.000",
           "device-timed | 1000.000 samples/sec (all 4 GPUs, max over ranks) | 1.0 ms/step",
           "Epoch 1: 72.0 seconds", " * Prec@1 20.000 Prec@5 60.000"]
    parsed = process_output.parse(log)
    assert parsed["top1"] == [10.0, 20.0] and parsed["epoch_hours"] == [0.01, 0.02]
    tbl = process_output.table(parsed).splitlines()
    assert tbl[0].startswith("Epoch#") and tbl[2].split("\t")[1] == "0.030"


def test_driver_yaml_to_command():
    conf = driver.load("configs/driver/resnet50_8hybrid.yml")
    cmd = driver.build_command(conf)
    assert "--nproc-per-node=8" in cmd and "resnet50" in cmd and "pipedream" in cmd
    assert driver.arch_of("models.vgg16.gpus=4") == "vgg16"
    with pytest.raises(ValueError):
        import yaml, tempfile
        with tempfile.NamedTemporaryFile("w", suffix=".yml", delete=False) as f:
            yaml.safe_dump({"module": "x"}, f)
        driver.load(f.name)


def test_activation_logger_and_sparsity(tmp_path):
    m = Block()
    lg = ActivationAndGradientLogger(str(tmp_path))
    lg.hook_modules(m)
    m(torch.randn(2, 3, 8, 8)).sum().backward()
    lg.dump(0)
    lg.unhook_modules()
    acts = pickle.load(open(tmp_path / "0" / "activations.pkl", "rb"))
    assert "r" in acts and "fc" in acts
    sp = sparsity.measure_sparsity(str(tmp_path / "0" / "activations.pkl"))
    assert 0.2 < sp["r"][0.0] < 0.8            # ReLU output is about half zeros
    cr = sparsity.measure_compression_ratio(str(tmp_path / "0" / "activations.pkl"))
    assert 0.1 < cr["r"] < 1.0      # zero-run-length coding shrinks a ReLU output
    assert sparsity.extract_reduce_times(["allreduce[2] start=1.5 end=2.0"]) == [(2, 0.5)]


def test_network_summary_counts_parameters():
    m = build_model("resnet18", "cifar10")
    text, total, trainable = summary.summary(m)
    assert total == sum(p.numel() for p in m.parameters()) == trainable
    assert "Conv2d-1" in text and "Total params" in text


def test_launcher_passes_ranks_and_propagates_failure(tmp_path):
    import subprocess
    import sys
    script = tmp_path / "child.py"
    script.write_text(
        "import argparse, os, sys\n"
        "ap = argparse.ArgumentParser(); ap.add_argument('--rank', type=int)\n"
        "ap.add_argument('--local_rank', type=int); ap.add_argument('--fail', type=int, default=-1)\n"
        "a = ap.parse_args()\n"
        "assert int(os.environ['RANK']) == a.rank and os.environ['WORLD_SIZE'] == '3'\n"
        "open(os.path.join(os.path.dirname(__file__), 'r%d' % a.rank), 'w').write(str(a.local_rank))\n"
        "sys.exit(7 if a.rank == a.fail else 0)\n")
    cmd = [sys.executable, "-m", "ddlbench_b200.launch", "--nproc_per_node", "3", "--pid-file",
           str(tmp_path / "pids"), str(script)]
    assert subprocess.run(cmd, timeout=120).returncode == 0
    assert sorted(p.name for p in tmp_path.glob("r*")) == ["r0", "r1", "r2"]
    assert len(open(tmp_path / "pids").read().split()) == 3
    assert subprocess.run(cmd + ["--fail", "1"], timeout=120).returncode == 7
    assert subprocess.run([sys.executable, "-m", "ddlbench_b200.launch", "--terminate",
                           str(tmp_path / "pids")], timeout=60).returncode == 0


def test_threadsafe_queue_and_counter():
    import threading
    from ddlbench_b200.utils.threadsafe import Counter, Queue
    q, c = Queue(maxsize=2), Counter(4)

    def prod():
        for i in range(4):
            q.add(i)
            c.decrement()
    t = threading.Thread(target=prod)
    t.start()
    got = [q.remove(timeout=5) for _ in range(4)]
    c.wait(timeout=5)
    t.join()
    assert got == [0, 1, 2, 3] and c.count == 0 and len(q) == 0


def test_runtime_stats_format():
    from ddlbench_b200.parallel.pipedream.stats import RuntimeStats
    st = RuntimeStats(forward=True)
    st.sent([torch.zeros(10, dtype=torch.float32)])
    tok = st.begin_compute("cpu")
    st.end_compute(tok)
    lines = []
    st.print_stats(out=lines.append)
    assert lines[0] == "Forward Stats:" and "\t send_tensors_size 40.000 bytes" in lines
    st.reset_stats()
    assert st.stats["send_tensors_size"] == 0


def test_gnmt_lr_schedule_and_fp_optimizers():
    from ddlbench_b200.harness.lr import WarmupMultiStepLR
    from ddlbench_b200.utils.fp_optimizer import Fp16Optimizer, Fp32Optimizer, flat_grad_norm
    m = nn.Linear(4, 4)
    opt = torch.optim.SGD(m.parameters(), lr=1.0)
    sch = WarmupMultiStepLR(opt, iterations=100, warmup_steps=10, remain_steps=60, decay_interval=10,
                            decay_steps=3, decay_factor=0.5)
    lrs = []
    for _ in range(100):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        sch.step()
    assert abs(lrs[0] - 0.01) < 1e-6 and lrs[5] < lrs[9] < 1.0 + 1e-9 and lrs[30] == 1.0
    assert lrs[60] == 0.5 and lrs[70] == 0.25 and lrs[99] == 0.125
    # fp32 wrapper clips to the requested norm
    w = Fp32Optimizer(m, grad_clip=0.1)
    loss = (m(torch.ones(2, 4)) * 100).sum()
    norm = w.step(loss, torch.optim.SGD(m.parameters(), lr=0.0))
    assert norm > 0.1
    # bf16 model + fp32 master copy stays in sync
    mb = nn.Linear(4, 4).to(torch.bfloat16)
    w16 = Fp16Optimizer(mb, grad_clip=5.0)
    o16 = torch.optim.SGD(w16.fp32_params, lr=0.1)
    before = mb.weight.detach().clone()
    w16.step(mb(torch.ones(2, 4, dtype=torch.bfloat16)).float().sum(), o16)
    assert not torch.equal(before, mb.weight) and w16.master.dtype == torch.float32
    assert flat_grad_norm(mb.parameters()) == 0.0


def test_gnmt_partition_tools_and_log_table(tmp_path):
    from ddlbench_b200.models.gnmt import GNMT
    from ddlbench_b200.partition import gnmt_tools
    m = GNMT(vocab_size=64, hidden_size=16, num_layers=4)
    times = gnmt_tools.profile_gnmt(m, batch_size=2, seq_len=4, iters=1)
    assert set(times) == {name for name, _, _, _ in m.spec} and all(f >= 0 for f, _ in times.values())
    cuts, reps = gnmt_tools.partition(m, 4, times, batch_size=2, seq_len=4, bandwidth=1e12, straight=True)
    assert cuts[-1] == len(m.spec) and len(cuts) == len(reps) and sum(reps) <= 4
    stages = m.split(len(cuts), cuts)
    assert sum(len(s.spec) for s in stages) == len(m.spec)
    d = gnmt_tools.write_confs(str(tmp_path), "gnmt", m, cuts, reps)
    import json
    conf = json.load(open(os.path.join(d, "hybrid_conf.json")))
    assert len(conf["module_to_stage_map"]) == len(m.spec) + 1 and conf["cuts"] == cuts
    assert os.path.exists(os.path.join(d, "mp_conf.json")) and os.path.exists(os.path.join(d, "dp_conf.json"))
    parsed = process_output.parse_gnmt(["Epoch: [0][0/5]\tTime 0.1\tLoss 9.5000", "Epoch 0: 36.0 seconds",
                                        "Epoch 0: BLEU 12.50"])
    assert parsed["bleu"] == [12.5] and parsed["loss"] == [9.5]
    assert process_output.table_gnmt(parsed).splitlines()[1].split("\t")[2] == "12.50"


def test_imagenet_subset_tool(tmp_path):
    import subprocess
    import sys
    src = tmp_path / "src"
    for split in ("train", "val"):
        for c in ("n01", "n02", "n03"):
            (src / split / c).mkdir(parents=True)
            for i in range(3):
                (src / split / c / ("%d.JPEG" % i)).write_bytes(b"x")
    r = subprocess.run([sys.executable, "tools/imagenet_subset.py", str(src), str(tmp_path / "dst"),
                        "--classes", "2", "--per-class", "2"], capture_output=True, text=True)
    assert r.returncode == 0 and "copied 8 files of 2 classes" in r.stdout
    assert len(os.listdir(tmp_path / "dst" / "train")) == 2


def test_step_timeline_chrome_trace(tmp_path):
    import json
    from ddlbench_b200.engine import StageExecutor
    from ddlbench_b200.utils.timeline import StepTimeline
    m = build_model("resnet18", "cifar10")
    ex = StageExecutor.from_model(m, device="cpu", dtype=torch.float32)
    tl = StepTimeline(ex)
    with tl:
        outs, ctx = ex.forward([torch.randn(2, 3, 32, 32)])
        _, d = ex.loss(outs[0], torch.tensor([1, 2]))
        ex.backward(ctx, [d])
    assert ex.profiler is None
    spans = tl.spans()
    assert len(spans) == 2 * len(ex.plan.ops) and all(dur >= 0 for _, _, dur in spans)
    trace = json.load(open(tl.save(str(tmp_path / "t.json"))))
    names = [e["name"] for e in trace["traceEvents"] if e["ph"] == "X"]
    assert any(n.startswith("fwd conv_bn_act") for n in names) and any(n.startswith("bwd ") for n in names)
    assert "ms" in tl.table(5).splitlines()[0]


def test_roofline_cost_model():
    from ddlbench_b200.utils.cost import model_cost, roofline
    m = build_model("resnet50", "imagenet")
    c = model_cost(m, 2)
    assert abs(c["forward_flops"] / 2 / 2 - 4.09e9) < 0.05e9          # 4.09 GMACs per image
    r = roofline(m, 32, 4.64)
    assert r["step_gflop"] > 700 and "of " in r["tensor_frac"] and r["peaks"]["hbm_gbs"] > 1000
    assert m.training                                                 # mode restored


def test_driver_multi_host_and_translation_commands():
    two = driver.build_node_commands(driver.load("configs/driver/resnet50_16hybrid_2nodes.yml"))
    assert [h for h, _ in two] == ["node0", "node1"] and all(c[0] == "ssh" for _, c in two)
    assert "--node_rank 1" in two[1][1][-1] and "--master_addr node0" in two[1][1][-1]
    assert "-g 8" in two[0][1][-1] and "-n 2" in two[0][1][-1]
    (host, cmd), = driver.build_node_commands(driver.load("configs/driver/gnmt_4pipedream.yml"))
    assert "ddlbench_b200.harness.translation" in cmd and cmd[cmd.index("--stages") + 1] == "4"


def test_driver_docker_mode_wraps_each_node_in_the_configured_container(tmp_path, capsys):
    """``--docker`` / ``--launch_single_container``: one container per node (the reference's
    ``driver.py:276-318``), repository + data + log directories mounted."""
    conf = driver.load("configs/driver/resnet50_16hybrid_2nodes.yml")
    two = driver.build_node_commands(conf, docker=True, mounts=["/scratch"])
    for i, (host, cmd) in enumerate(two):
        assert cmd[0] == "ssh" and host == "node%d" % i
        inner = cmd[-1]
        assert inner.startswith("docker run") and "--gpus all" in inner and "--net=host" in inner
        assert str(conf["container"]) in inner and "-v /scratch:/scratch" in inner
        assert "ddlbench_b200.launch" in inner and "--node_rank %d" % i in inner
    one_conf = driver.load("configs/driver/vgg16_4pipedream.yml")
    (host, cmd), = driver.build_node_commands(one_conf, docker=True)
    assert cmd[:2] == ["docker", "run"] and cmd[-2] == "-c" and "torch.distributed.run" in cmd[-1]
    assert cmd[cmd.index("/bin/bash") - 1] == str(one_conf["container"])
    # the CLI prints the same thing with --dry_run
    assert driver.main(["--config_file", "configs/driver/vgg16_4pipedream.yml", "--dry_run", "--docker"]) == 0
    assert capsys.readouterr().out.startswith("docker run")


def test_launcher_terminate_stops_exactly_the_recorded_pids(tmp_path):
    """``launch.py --terminate <pidfile>`` signals the recorded children only (the reference's
    ``terminate_runtime.py`` runs ``pkill -f`` on every machine)."""
    import signal
    import subprocess
    import sys
    import time
    script = tmp_path / "sleeper.py"
    script.write_text("import time, sys\ntime.sleep(120)\n")
    pidfile = tmp_path / "pids"
    launcher = subprocess.Popen([sys.executable, "-m", "ddlbench_b200.launch", "--nproc_per_node", "2",
                                 "--no-rank-args", "--pid-file", str(pidfile), str(script)])
    bystander = subprocess.Popen([sys.executable, str(script)])     # same command line, not ours
    try:
        for _ in range(100):
            if pidfile.exists() and len(pidfile.read_text().split()) == 2:
                break
            time.sleep(0.1)
        pids = [int(p) for p in pidfile.read_text().split()]
        assert len(pids) == 2 and bystander.pid not in pids
        r = subprocess.run([sys.executable, "-m", "ddlbench_b200.launch", "--terminate", str(pidfile)],
                           capture_output=True, text=True, timeout=60)
        assert r.returncode == 0 and "Done Terminating" in r.stdout
        assert launcher.wait(timeout=30) != 0            # its children were terminated
        assert bystander.poll() is None                  # an unrelated process with the same name lives on
    finally:
        bystander.kill()
        bystander.wait()
        if launcher.poll() is None:
            launcher.kill()

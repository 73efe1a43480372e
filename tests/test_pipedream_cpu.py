"""PipeDream runtime on the CPU/Gloo fake transport: schedules, routing, weight-stashing
semantics, and multi-process end-to-end equivalences."""
import glob
import os
import subprocess
import sys

import pytest
import torch

from ddlbench_b200.engine import StageExecutor
from ddlbench_b200.models import build_model
from ddlbench_b200.parallel.pipedream.config import PipelineConfig, RankTopology
from ddlbench_b200.parallel.pipedream.schedule import (MessagingCursor, one_f_one_b, validate)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PORT = [29600]


def _launch(reps, tmp_path, extra_env=None, iters=6, timeout=420):
    n = sum(reps)
    env = dict(os.environ)
    env.update({"PD_REPS": ",".join(map(str, reps)), "PD_OUT": str(tmp_path), "PD_CPU": "1",
                "PD_ITERS": str(iters), "OMP_NUM_THREADS": "2"})
    env.update(extra_env or {})
    _PORT[0] += 1
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(_PORT[0]),
           os.path.join(ROOT, "tests", "pipedream_worker.py")]
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return [torch.load(f) for f in sorted(glob.glob(os.path.join(str(tmp_path), "rank*.pt")))]


def test_warmup_formula_and_confs():
    cfg = PipelineConfig.straight(4)
    assert [cfg.warmup(s) for s in range(4)] == [3, 2, 1, 0]
    assert cfg.module_to_stage_map == [0, 1, 2, 3, 3]
    h = PipelineConfig.hybrid(PipelineConfig.parse_stage_to_num_ranks("0:3,1:1"))
    assert h.stage_to_rank_map == {0: [0, 1, 2], 1: [3]} and [h.warmup(0), h.warmup(1)] == [1, 0]
    assert h.lr_scale(1) == pytest.approx(1 / 3)
    d = PipelineConfig.data_parallel(3, 4)
    assert d.warmup(0) == 0 and d.num_stages == 1
    rt = PipelineConfig.from_json(h.to_dict())
    assert rt.stage_to_rank_map == h.stage_to_rank_map
    depth = PipelineConfig.from_json({"module_to_stage_map": [0, 1, 1],
                                      "stage_to_rank_map": {"0": [0], "1": [1]},
                                      "stage_to_depth_map": {"0": 5, "1": 0}})
    assert depth.warmup(0) == 5
    t = RankTopology.of(h, 1)
    assert (t.stage, t.rank_in_stage, t.next_ranks, t.prev_ranks) == (0, 1, [3], [])


def test_one_f_one_b_order():
    seq = one_f_one_b(5, 2)
    assert seq[:3] == [("F", 0), ("F", 1), ("F", 2)] and seq[3] == ("B", 0)
    assert [i for op, i in seq if op == "B"] == list(range(5))
    assert one_f_one_b(3, 0) == [("F", 0), ("B", 0), ("F", 1), ("B", 1), ("F", 2), ("B", 2)]


def test_messaging_cursor_when_previous_stage_is_smaller():
    # 4 replicas fed by 2: the case the reference's `>` -> `>=` fix is about
    seqs = [[MessagingCursor(r, 4, 2).take() for _ in range(1)] for r in range(4)]
    assert seqs == [[0], [1], [1], [1]]
    c = MessagingCursor(0, 2, 4)
    assert [c.take() for _ in range(4)] == [0, 2, 1, 3]


@pytest.mark.parametrize("reps", [[1, 1, 1, 1], [2, 1], [1, 2], [3, 1], [2, 2, 1], [4, 2, 1, 1],
                                  [2, 4], [7, 3, 6], [1, 1, 2, 4], [4, 2, 2]])
def test_routing_is_deadlock_free(reps):
    validate(PipelineConfig.hybrid(reps), 48)


def test_weight_stashing_semantics():
    """Idea of pipedream-fork/runtime/tests/backprop/sgd_with_stashing.py: forward three
    minibatches with the same weights, then interleave backward/step.  With V versions the
    first V backwards see the weights their forward saw (input-gradients equal those of a
    fresh model); later ones see stepped weights."""
    torch.manual_seed(0)
    m = build_model("vgg11", "mnist")
    x = torch.randn(4, 1, 28, 28)
    t = torch.randint(0, 10, (4,))

    def input_grads(versions):
        ex = StageExecutor.from_model(m, device="cpu", dtype=torch.float32, versions=versions,
                                      lr=0.5, momentum=0.0, first_stage=False)
        ctxs = []
        for _ in range(3):
            o, c = ex.forward([x], version=ex.store.latest, need_input_grad=True)
            s, d = ex.loss(o[0], t)
            ctxs.append((c, d))
        grads = []
        for c, d in ctxs:
            # backward against the version recorded at forward time — unless it was recycled
            if (ex.store.latest - c.version) >= versions:
                c.version = ex.store.latest
            grads.append(ex.backward(c, [d])[0].clone())
            ex.store.sgd_step(out_version=ex.store.next_version())
            ex.store.advance_version()
        return grads

    g1, g2, g3 = input_grads(1), input_grads(2), input_grads(3)
    assert not torch.allclose(g1[0], g1[1])                 # V=1: second backward sees new weights
    assert torch.allclose(g2[0], g2[1]) and not torch.allclose(g2[1], g2[2])
    assert torch.allclose(g3[0], g3[1]) and torch.allclose(g3[1], g3[2])
    assert torch.allclose(g1[0], g3[0])


def test_no_input_pipelining_equals_sequential_sgd(tmp_path):
    """warm-up 0, one weight version == plain model-parallel == single-process SGD."""
    res = _launch([1, 1], tmp_path, {"PD_NOPIPE": "1"}, iters=4)
    torch.manual_seed(0)
    model = build_model("resnet18", "cifar10")
    ex = StageExecutor.from_model(model, device="cpu", dtype=torch.float32, lr=0.05, momentum=0.9,
                                  weight_decay=1e-4)
    g = torch.Generator().manual_seed(100)
    losses = []
    for _ in range(4):
        x, y = torch.randn(8, 3, 32, 32, generator=g), torch.randint(0, 10, (8,), generator=g)
        o, c = ex.forward([x])
        s, d = ex.loss(o[0], y)
        ex.backward(c, [d])
        ex.store.sgd_step()
        losses.append(float(s[0]))
    ref = ex.state_dict()
    assert all(r["warmup"] == 0 and r["versions"] == 1 for r in res)
    pipe_losses = [r for r in res if r["losses"]][0]["losses"]
    assert pipe_losses == pytest.approx(losses, rel=5e-3, abs=5e-3)  # fp32 summation order differs with thread count
    for r in res:
        for k, v in r["state"].items():
            assert torch.allclose(v.float(), ref[k].float(), atol=5e-3, rtol=5e-2), k


def test_1f1b_two_stage_runs_with_stashing(tmp_path):
    res = _launch([1, 1], tmp_path, iters=6)
    s0 = [r for r in res if r["stage"] == 0][0]
    s1 = [r for r in res if r["stage"] == 1][0]
    assert (s0["warmup"], s0["versions"]) == (1, 2) and (s1["warmup"], s1["versions"]) == (0, 1)
    assert len(s1["losses"]) == 6 and all(l == l for l in s1["losses"])
    assert s0["bytes_sent"] > 0 and s1["bytes_sent"] > 0


def test_hybrid_replicated_first_stage(tmp_path):
    res = _launch([2, 1], tmp_path, iters=4)
    reps = [r for r in res if r["stage"] == 0]
    assert len(reps) == 2
    for k in reps[0]["state"]:
        if "running_" in k:
            continue   # BatchNorm statistics are local to a replica (as with DDP in the reference)
        assert torch.allclose(reps[0]["state"][k], reps[1]["state"][k], atol=1e-6), k
    last = [r for r in res if r["stage"] == 1][0]
    assert len(last["losses"]) == 8   # 4 per first-stage replica x 2 replicas


def test_macrobatch_mode(tmp_path):
    """Macrobatch (2BW-like): <= 2 versions and EVERY stage averages the minibatch gradients of its
    own update interval (reference optimizer.py:140-142 divides p.grad by update_interval in each
    stage's optimizer), not just the last one through the loss."""
    res = _launch([1, 1, 1], tmp_path, {"PD_MACRO": "1"}, iters=6)
    by_stage = {r["stage"]: r for r in res}
    assert by_stage[0]["versions"] <= 2
    for s, r in by_stage.items():
        assert r["update_interval"] == r["warmup"] + 1
        assert r["grad_scale"] == pytest.approx(1.0 / r["update_interval"])
    assert by_stage[0]["update_interval"] == 3 and by_stage[2]["update_interval"] == 1


def test_accumulated_gradients_are_averaged_by_grad_scale():
    """What the macrobatch scaling relies on: k accumulated minibatch gradients with
    grad_scale = 1/k give the SGD step of their mean."""
    torch.manual_seed(0)
    m = build_model("vgg11", "mnist")
    xs = [torch.randn(4, 1, 28, 28) for _ in range(3)]
    ts = [torch.randint(0, 10, (4,)) for _ in range(3)]

    def make():
        return StageExecutor.from_model(m, device="cpu", dtype=torch.float32, lr=0.5, momentum=0.0)

    a = make()
    a.store.set_grad_scale(1.0 / 3)
    for i, (x, t) in enumerate(zip(xs, ts)):
        o, c = a.forward([x])
        _, d = a.loss(o[0], t)
        a.backward(c, [d], accumulate=i > 0, final=i == 2)
    a.store.sgd_step()
    b = make()
    g = torch.zeros_like(b.store.g_flat)
    for x, t in zip(xs, ts):
        o, c = b.forward([x])
        _, d = b.loss(o[0], t)
        b.backward(c, [d])
        g += b.store.g_flat
    b.store.g_flat.copy_(g / 3)
    b.store.sgd_step()
    assert torch.allclose(a.store.master, b.store.master, atol=1e-6)


def test_stage_graph_slots_match_eager(tmp_path):
    """DDLB_PIPE_GRAPHS: the per-version-slot replay logic (static inputs, busy slots during the
    warm-up forwards, static gradient buffers) must reproduce the eager 1F1B run exactly.  On the
    CPU the graph object is a stand-in that re-runs the captured body."""
    a = tmp_path / "eager"
    b = tmp_path / "graphs"
    a.mkdir(), b.mkdir()
    eager = _launch([1, 1, 1], a, iters=10)
    graphed = _launch([1, 1, 1], b, {"DDLB_PIPE_GRAPHS": "1", "DDLB_PIPE_GRAPH_AFTER": "2"}, iters=10)
    for e, g in zip(eager, graphed):
        assert e["stage"] == g["stage"] and e["graph_replays"] == 0
        assert g["graph_replays"] > 0 and g["graph_slots"] == g["versions"]
        assert e["losses"] == g["losses"]
        for k, v in e["state"].items():
            assert torch.equal(v, g["state"][k]), k


def _cli_pipedream(tmp_path, port, *extra):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", str(port), "-m", "ddlbench_b200.cli",
           "-b", "mnist", "-f", "pipedream", "-g", "2", "-m", "resnet18", "--worker", "--device", "cpu",
           "--steps", "3", "--eval-steps", "1", "--batch-size", "8", "--stages", "2x1", *extra]
    env = dict(os.environ, OMP_NUM_THREADS="2", DDLB_STAGE_BALANCE="size",
               PYTHONPATH=ROOT + os.pathsep + os.environ.get("PYTHONPATH", ""))
    p = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600, cwd=str(tmp_path))
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    return p.stdout


def test_resume_continues_from_checkpoint(tmp_path):
    """--resume loads weights + fp32 master + momentum + BatchNorm buffers + epoch
    (main_with_runtime.py:240-262 --resume): one 2-epoch run == 1 epoch, stop, resume to epoch 2."""
    a, b = tmp_path / "A", tmp_path / "B"
    _cli_pipedream(tmp_path, 29811, "--epochs", "2", "--checkpoint-dir", str(a))
    _cli_pipedream(tmp_path, 29812, "--epochs", "1", "--checkpoint-dir", str(b))
    first = torch.load(b / "checkpoint.0.pth.tar")
    assert first["epoch"] == 1
    out = _cli_pipedream(tmp_path, 29813, "--epochs", "2", "--resume", str(b), "--checkpoint-dir", str(b))
    assert "loaded checkpoint" in out and "Epoch 0:" not in out and "Epoch 1:" in out
    for stage in (0, 1):
        x = torch.load(a / ("checkpoint.%d.pth.tar" % stage))
        y = torch.load(b / ("checkpoint.%d.pth.tar" % stage))
        assert x["epoch"] == y["epoch"] == 2
        assert torch.equal(x["master"], y["master"]) and torch.equal(x["momentum"], y["momentum"])
        for k, v in x["state_dict"].items():
            assert torch.equal(v, y["state_dict"][k]), k
    # the weights moved between the epoch-1 and the epoch-2 checkpoint (the run really continued)
    assert not torch.equal(first["master"], torch.load(b / "checkpoint.0.pth.tar")["master"])


def test_sharded_optimizer_state_round_trip():
    """gather_full_state / load_full_state: what a checkpoint of a fused-replica (sharded) stage
    goes through — slices of every bucket from every rank and back."""
    from ddlbench_b200.engine.store import ParamSpec, ParamStore
    specs = [ParamSpec("a", (5, 7), torch.randn(5, 7), False, 0), ParamSpec("b", (33,), torch.randn(33), False, 1),
             ParamSpec("c", (4, 3, 3, 3), torch.randn(4, 3, 3, 3), True, 2)]
    full = ParamStore(specs, "cpu", torch.float32, bucket_elems=40)
    full.mom.copy_(torch.randn_like(full.mom))
    world = 2
    shards = [ParamStore(specs, "cpu", torch.float32, bucket_elems=40, world=world, rank=r, shard_optimizer=True)
              for r in range(world)]
    layout_total = shards[0].layout.total
    m = torch.randn(layout_total)
    v = torch.randn(layout_total)
    for s in shards:
        s.load_full_state(m, v)
    # reassemble by hand what all_gather would deliver
    rebuilt_m, rebuilt_v = torch.zeros(layout_total), torch.zeros(layout_total)
    for bkt, (mo, ln) in zip(shards[0].layout.buckets, shards[0].master_slices()):
        for r, s in enumerate(shards):
            rebuilt_m[bkt.offset + r * ln: bkt.offset + (r + 1) * ln] = s.master[mo:mo + ln]
            rebuilt_v[bkt.offset + r * ln: bkt.offset + (r + 1) * ln] = s.mom[mo:mo + ln]
    assert torch.equal(rebuilt_m, m) and torch.equal(rebuilt_v, v)
    assert torch.equal(shards[0].w_flat[0], m)          # every replica holds the full weights


def test_parallel_config_schema_reserves_tp_cp():
    """tp / cp are reserved in the conf schema; only 1 is accepted today."""
    base = {"module_to_stage_map": [0, 1, 1], "stage_to_rank_map": {"0": [0], "1": [1]}}
    cfg = PipelineConfig.from_json(dict(base, parallel={"tp": 1, "cp": 1}))
    assert (cfg.tp, cfg.cp) == (1, 1) and "parallel" not in cfg.to_dict()
    for bad in ({"tp": 2}, {"cp": 4}):
        with pytest.raises(NotImplementedError):
            PipelineConfig.from_json(dict(base, parallel=bad))


def test_alexnet_and_driver_configs():
    """AlexNet is in the zoo (reference: alexnet_{4,16}{dp,pipedream}.yml) and every driver YAML of
    the reference has a counterpart that builds a launch command."""
    from ddlbench_b200 import driver
    m = build_model("alexnet", "imagenet")
    assert sum(p.numel() for p in m.parameters()) == 61100840
    names = {"alexnet_16dp", "alexnet_16pipedream", "alexnet_4dp", "alexnet_4pipedream", "resnet50_16dp",
             "vgg16_16dp", "vgg16_16pipedream", "vgg16_4dp", "vgg16_4pipedream", "gnmt_12dp", "gnmt_12pipedream",
             "gnmt_4dp", "gnmt_4pipedream", "gnmt_large_16dp", "gnmt_large_16pipedream", "gnmt_large_4dp",
             "gnmt_large_4pipedream"}
    have = {f[:-4] for f in os.listdir(os.path.join(ROOT, "configs", "driver")) if f.endswith(".yml")}
    assert names <= have, names - have
    for n in sorted(names):
        conf = driver.load(os.path.join(ROOT, "configs", "driver", n + ".yml"))
        nodes = driver.build_node_commands(conf)
        assert nodes and all(cmd for _, cmd in nodes)
    dp = driver.build_command(driver.load(os.path.join(ROOT, "configs", "driver", "alexnet_4dp.yml")))
    assert "--stages" in dp and dp[dp.index("--stages") + 1] == "1x4" and "alexnet" in dp


def test_loss_scale_is_transparent(tmp_path):
    """--loss_scale S: gradients carry S through the whole pipeline and every stage's optimizer
    divides it out — the update equals the unscaled run (fp32 on the CPU: up to rounding)."""
    a, b = tmp_path / "A", tmp_path / "B"
    _cli_pipedream(tmp_path, 29821, "--epochs", "1", "--checkpoint-dir", str(a))
    _cli_pipedream(tmp_path, 29822, "--epochs", "1", "--checkpoint-dir", str(b), "--loss_scale", "128")
    for stage in (0, 1):
        x = torch.load(a / ("checkpoint.%d.pth.tar" % stage))
        y = torch.load(b / ("checkpoint.%d.pth.tar" % stage))
        assert float(y["hyper"][3]) == pytest.approx(float(x["hyper"][3]) / 128.0)
        assert torch.allclose(x["master"], y["master"], rtol=1e-4, atol=1e-6)

"""Reference-compatible log lines — kept grep-compatible with the
reference's ``slurm.out`` so existing post-processing keeps working."""
from __future__ import annotations

from typing import Tuple

import torch


def creating_model(arch: str) -> str:
    # benchmark/imagenet/imagenet_pytorch.py:62
    return "=> creating model '{}'".format(arch)


def memory_gb(device=None) -> Tuple[float, float, float]:
    """(peak allocated, peak reserved, total) in GB — imagenet_pytorch.py:165-172."""
    if torch.cuda.is_available():
        dev = torch.cuda.current_device() if device is None else device
        stats = torch.cuda.memory_stats(device=dev)
        total = torch.cuda.get_device_properties(dev).total_memory
        return (stats.get("allocated_bytes.all.peak", 0) / 1e9,
                stats.get("reserved_bytes.all.peak", 0) / 1e9, float(total) / 1e9)
    return (0.0, 0.0, 0.0)


def train_progress(epoch: int, epochs: int, percent: float, throughput: float, device=None) -> str:
    a, r, t = memory_gb(device)
    return ("train | %d/%d epoch (%d%%) | %.3f samples/sec (estimated) | mem (GB): %.3f (%.3f) / %.3f"
            % (epoch, epochs, percent, throughput, a, r, t))


def valid_progress(percent: float, throughput: float) -> str:
    return "valid | %d%% | %.3f samples/sec (estimated)" % (percent, throughput)


def epoch_summary(epoch: int, epochs: int, train_loss: float, throughput: float,
                  valid_loss: float, valid_acc: float) -> str:
    return ("%d/%d epoch | train loss:%.3f %.3f samples/sec | valid loss:%.3f accuracy:%.3f"
            % (epoch, epochs, train_loss, throughput, valid_loss, valid_acc))


def final_summary(valid_acc: float, throughput: float, sec_per_epoch: float) -> str:
    return ("valid accuracy: %.4f | %.3f samples/sec, %.3f sec/epoch (average)"
            % (valid_acc, throughput, sec_per_epoch))


def test_set(loss: float, acc_percent: float) -> str:
    # benchmark/imagenet/imagenet_horovod.py:252-253
    return "\nTEST SET: Average loss: {:.4f}, Accuracy: {:.2f}%\n".format(loss, acc_percent)


def device_timed(throughput: float, ms_per_step: float, world: int) -> str:
    """Extra line (not in the reference): the device-timed, max-over-ranks number that
    BASELINE.json defines as the metric."""
    return ("device-timed | %.3f samples/sec (all %d GPUs, max over ranks) | %.3f ms/step"
            % (throughput, world, ms_per_step))


# PipeDream harness lines (pipedream-fork/runtime/image_classification/main_with_runtime.py:458-500)
def pd_train(epoch, i, n, t_val, t_avg, epoch_hr, mem_a, mem_c, loss_val, loss_avg, p1, p1a, p5,
             p5a) -> str:
    return ("Epoch: [{0}][{1}/{2}]\tTime: {3:.3f} ({4:.3f})\tEpoch time [hr]: {5:.3f} ({5:.3f})\t"
            "Memory: {6:.3f} ({7:.3f})\tLoss: {8:.4f} ({9:.4f})\tPrec@1: {10:.3f} ({11:.3f})\t"
            "Prec@5: {12:.3f} ({13:.3f})").format(epoch, i, n, t_val, t_avg, epoch_hr, mem_a,
                                                  mem_c, loss_val, loss_avg, p1, p1a, p5, p5a)


def pd_train_nonlast(epoch, i, n, mem_a, mem_c) -> str:
    return "Epoch: [{0}][{1}/{2}]\tMemory: {3:.3f} ({4:.3f})".format(epoch, i, n, mem_a, mem_c)


def pd_test(epoch, i, n, t_val, t_avg, mem_a, mem_c, loss_val, loss_avg, p1, p1a, p5, p5a) -> str:
    return ("Test: [{0}][{1}/{2}]\tTime: {3:.3f} ({4:.3f})\tMemory: {5:.3f} ({6:.3f})\t"
            "Loss: {7:.4f} ({8:.4f})\tPrec@1: {9:.3f} ({10:.3f})\tPrec@5: {11:.3f} ({12:.3f})"
            ).format(epoch, i, n, t_val, t_avg, mem_a, mem_c, loss_val, loss_avg, p1, p1a, p5, p5a)


def pd_prec(p1: float, p5: float) -> str:
    return " * Prec@1 {:.3f} Prec@5 {:.3f}".format(p1, p5)


def pd_epoch_time(epoch: int, seconds: float) -> str:
    return "Epoch %d: %.3f seconds" % (epoch, seconds)


def pd_lr(epoch: int, step: int, lr: float) -> str:
    return "Epoch: %d Step %d \tLearning rate: %f" % (epoch, step, lr)

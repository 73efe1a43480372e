from .loop import StepTimer, evaluate, max_over_ranks, run, timed_loop
from .trainer import Trainer

__all__ = ["StepTimer", "evaluate", "max_over_ranks", "run", "timed_loop", "Trainer"]

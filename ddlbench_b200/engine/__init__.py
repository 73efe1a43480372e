"""Static (autograd-free) fused-op execution engine."""
from .executor import FwdContext, StageExecutor
from .plan import Plan, compile_plan
from .store import FlatLayout, ParamSpec, ParamStore

__all__ = ["FwdContext", "StageExecutor", "Plan", "compile_plan", "FlatLayout", "ParamSpec",
           "ParamStore"]

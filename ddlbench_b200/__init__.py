"""ddlbench_b200 — a Blackwell-native distributed deep-learning benchmark framework
with the capabilities of sara-nl/DDLBench (single GPU, Horovod-style data parallel,
GPipe, PipeDream 1F1B) — see DESIGN.md."""
__version__ = "0.1.0"

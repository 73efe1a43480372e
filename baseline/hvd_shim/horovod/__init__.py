# Minimal stand-in for the `horovod` package (NOT part of ddlbench_b200; see torch/__init__.py).

# Convenience targets (everything is plain python underneath)
PY ?= python

.PHONY: build test test-gpu bench bench8 sass clean

build:            ## compile the sm_100a extension in-tree
	$(PY) -m ddlbench_b200.ops.build

test:             ## CPU test-suite (Gloo for multi-process paths)
	$(PY) -m pytest tests -x -q -m "not gpu"

test-gpu:         ## on a B200
	$(PY) -m pytest tests -x -q -m gpu

bench:            ## headline benchmark, 1 GPU
	$(PY) bench.py --gpus 1 --steps 50 --warmup 5

bench8:           ## headline benchmark, 8 GPUs of one node
	$(PY) -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
	    --master-port 29500 bench.py --gpus 8 --steps 50 --warmup 5

sass:             ## per-kernel SASS evidence -> profiles/sass/
	$(PY) tools/sass_evidence.py

clean:
	rm -rf build ddlbench_b200/ops/_ddlb_ext*.so

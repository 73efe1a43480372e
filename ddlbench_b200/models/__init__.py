from .graph import Add, Flatten, GraphBuilder, GraphSegment, INPUT, LayerGraphModel, LayerNode
from .zoo import (ARCHS, BENCH_ARCHS, DATASETS, build_model, input_shape, num_classes)
from . import zoo

__all__ = ["Add", "Flatten", "GraphBuilder", "GraphSegment", "INPUT", "LayerGraphModel",
           "LayerNode", "ARCHS", "BENCH_ARCHS", "DATASETS", "build_model", "input_shape",
           "num_classes", "zoo"]

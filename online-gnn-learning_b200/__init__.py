"""ogl_b200 -- B200-native streaming GraphSAGE hot path (drop-in for online-gnn-learning's
`--backend pytorch --cuda`).  Host code is Python/PyTorch plumbing over hand-written sm_100a kernels
reached through the C ABI in include/ogl_b200.h; there is no CPU or eager fallback."""
from . import config                                  # noqa: F401
from ._lib import OGL_F32, OGL_BF16, OGL_TF32, OGL_FP16, OglError, LIB_PATH, kernel_launches          # noqa: F401
from . import _native as native                       # noqa: F401
from . import utils, sampling, parallel, inference    # noqa: F401
from .utils import Lib_supported, init                # noqa: F401
from .graph.dynamic_graph import DynamicGraph         # noqa: F401
from .graph.dynamic_graph_edge import DynamicGraphEdge                     # noqa: F401
from .graph.dynamic_graph_vertex import DynamicGraphVertex, ParentGraph    # noqa: F401
from .graph.device_graph import DeviceGraph           # noqa: F401
from .graph.train_test_graph import TrainTestGraph    # noqa: F401
from .prioritized_replay.replay_buffer import PrioritizedReplayBuffer      # noqa: F401
from .prioritized_replay.segment_tree import SumSegmentTree                # noqa: F401
from .prioritized_replay.generate_priority import LossPriority, TrendPriority, HybridPriority   # noqa: F401
from .graphsage.pytorch.graphsage_dgl import GraphSAGE                      # noqa: F401

__all__ = ["config", "native", "utils", "sampling", "parallel", "inference", "init", "Lib_supported", "DynamicGraph", "DynamicGraphEdge",
           "DynamicGraphVertex", "ParentGraph", "DeviceGraph", "TrainTestGraph", "PrioritizedReplayBuffer",
           "SumSegmentTree", "LossPriority", "TrendPriority", "HybridPriority", "GraphSAGE", "kernel_launches"]

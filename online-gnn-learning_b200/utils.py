"""Backend registry + tensor helpers: the drop-in seam of the reference's train/utils.py.

`init(library, GPU, GPU_ID)` returns the same 6-tuple the reference's driver unpacks
(train/__main__.py:99 <- train/utils.py:16-38), bound to the B200 kernels.  Only the
PyTorch/CUDA backend exists here; the TF backends are out of scope.
"""
import enum
import numpy as np
import torch


class Lib_supported(enum.Enum):
    PYTORCH = 1
    TF = 2
    TF_STATIC = 3


LIB = Lib_supported.PYTORCH
_GPU = True
_GPU_ID = -1


def init(library=Lib_supported.PYTORCH, GPU=True, GPU_ID=-1):
    """-> (GraphSAGE, RandomT, PrioritizedT, NoRehT, FullT, activation)."""
    global LIB, _GPU, _GPU_ID
    if library != Lib_supported.PYTORCH:
        raise NotImplementedError("ogl_b200 implements the pytorch backend only (reference utils.py:21-38)")
    if not GPU:
        raise RuntimeError("ogl_b200 is the `--backend pytorch --cuda` path: there is no CPU fallback")
    LIB, _GPU, _GPU_ID = library, GPU, GPU_ID
    if GPU_ID is not None and GPU_ID >= 0:
        torch.cuda.set_device(int(GPU_ID))      # the reference passes the bool here (utils.py:30-31)
    from .graphsage.pytorch.graphsage_dgl import GraphSAGE
    from .graphsage.pytorch.model import (RandomPytorchSupervisedGraphSage, PrioritizedPytorchSupervisedGraphSage,
                                          NoRehPytorchSupervisedGraphSage, FullPytorchSupervisedGraphSage)
    return (GraphSAGE, RandomPytorchSupervisedGraphSage, PrioritizedPytorchSupervisedGraphSage,
            NoRehPytorchSupervisedGraphSage, FullPytorchSupervisedGraphSage, torch.nn.functional.relu)


def to_nn_lib(data, GPU=True, dtype=None):
    """torch tensor from array-like; float64 -> float32 (reference utils.py:62-66)."""
    t = data.detach().clone() if isinstance(data, torch.Tensor) else torch.as_tensor(np.asarray(data))
    if dtype is not None:
        t = t.to(dtype)
    if t.dtype == torch.float64:
        t = t.float()
    return t.cuda() if GPU else t


def from_nn_lib_to_list(data):
    return data.tolist()


def from_nn_lib_to_set(data):
    return set(data.tolist())


def from_nn_lib_to_numpy(data):
    return data.detach().cpu().numpy()


def from_nn_get_python_value(tensor):
    return tensor.item()


def get_context():
    return torch.device("cuda", torch.cuda.current_device())


def index_tensor(tensor, indices):
    if isinstance(indices, (list, tuple)):
        indices = torch.as_tensor(np.asarray(indices, dtype=np.int64), device=tensor.device)
    elif isinstance(indices, np.ndarray):
        indices = torch.as_tensor(indices.astype(np.int64), device=tensor.device)
    return tensor[indices]


class sparse1d:
    """original id -> subgraph id map, indexable by scalars or numpy arrays (the reference wraps
    a 1xV scipy CSC, utils.py:132-142; a dense int64 vector does the same job)."""

    def __init__(self, size, fill=0):
        self.vec = np.full(int(size), fill, dtype=np.int64)

    def __getitem__(self, items):
        if hasattr(items, "__len__") and not isinstance(items, str):
            return self.vec[np.asarray(items, dtype=np.int64)]
        return int(self.vec[int(items)])

    def __setitem__(self, keys, items):
        self.vec[keys] = items

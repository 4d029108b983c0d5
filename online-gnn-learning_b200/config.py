"""Process-wide settings of the B200 backend."""
from ._lib import OGL_F32, OGL_BF16, OGL_TF32, OGL_FP16

_STATE = {"precision": "tf32", "seed": 1, "faithful": True}      # default: the tensor-core mode with fp32 range that meets rtol 1e-3


def set_precision(name):
    """'bf16' (tcgen05 kind::f16 tensor-core path, bf16 storage: fastest; a few 1e-3 of the tensor scale away from the reference's fp32
    path), 'tf32' (tcgen05 kind::tf32 on fp32 storage: the tensor-core mode that meets rtol 1e-3 against the fp32 path), 'fp16'
    (tcgen05 kind::f16 on fp16 storage with static loss scaling: TF32's ten mantissa bits at bf16's speed, for standardised
    features -- stored values must stay below 65504) or 'fp32' (SIMT FFMA path, rtol 1e-5 vs the fp32 oracle)."""
    assert name in ("bf16", "tf32", "fp16", "fp32")
    _STATE["precision"] = name


def precision():
    return _STATE["precision"]


def mode():
    return {"bf16": OGL_BF16, "tf32": OGL_TF32, "fp16": OGL_FP16}.get(_STATE["precision"], OGL_F32)


def set_seed(seed):
    """Philox key of the neighbour sampler / RBR draws (the reference never seeds DGL)."""
    _STATE["seed"] = int(seed)


def seed():
    return _STATE["seed"]


def set_faithful(flag):
    """faithful=True reproduces the reference's vertex choosers literally (Python `random`
    shuffles, numpy train/test split, PBR falling back to uniform draws: SURVEY 8(a) a10/a11);
    faithful=False uses the on-GPU counter-RNG draws and real proportional PBR sampling."""
    _STATE["faithful"] = bool(flag)


def faithful():
    return _STATE["faithful"]

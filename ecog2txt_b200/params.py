"""Host-side parameter initialisation and (de)serialisation.  Names / shapes follow the TF-checkpoint
convention parsed by MultiSubjectTrainer.recover_model_sizes (/root/reference/ecog2txt/trainers.py:444-554)."""
from __future__ import annotations

import math
import os
import re
from typing import Dict, Optional

import numpy as np

from . import _lib as L


def glorot_init(shapes: Dict[str, tuple], seed: int = 1) -> Dict[str, np.ndarray]:
    """Glorot-uniform matrices (TF1 default initializer), zero biases; draws follow tensor order."""
    rs = np.random.RandomState(seed)
    out = {}
    for name, shape in shapes.items():
        if len(shape) == 1:
            out[name] = np.zeros(shape, np.float32)
            continue
        fan = (shape[1] * shape[2] + shape[3]) if len(shape) == 4 else (shape[0] + shape[1])
        lim = math.sqrt(6.0 / fan)
        out[name] = rs.uniform(-lim, lim, size=shape).astype(np.float32)
    return out


def init_engine(engine, seed: int = 1):
    shapes = {k: s for k, (s, _) in engine.tensors().items()}
    engine.set_all(glorot_init(shapes, seed))


EMA_SUFFIX = "/ExponentialMovingAverage"  # trainers.py:466-468


def save_checkpoint(engine, checkpoint_path: str, epoch: int) -> str:
    """<checkpoint_path>-<epoch>.index (manifest, discovered by restore_epoch: trainers.py:240-252)
    + <checkpoint_path>-<epoch>.npz holding variables, EMA shadows and Adam slots under TF names."""
    arrays = {}
    for name in engine.tensors():
        arrays[name] = engine.get(name, L.VALUE)
        arrays[name + EMA_SUFFIX] = engine.get(name, L.EMA)
        arrays[name + "/Adam"] = engine.get(name, L.ADAM_M)
        arrays[name + "/Adam_1"] = engine.get(name, L.ADAM_V)
    arrays["global_step"] = np.asarray(engine.step, np.int64)
    base = f"{checkpoint_path}-{epoch}"
    os.makedirs(os.path.dirname(os.path.abspath(base)), exist_ok=True)
    np.savez(base + ".npz", **arrays)
    with open(base + ".index", "w") as f:
        for k, v in arrays.items():
            f.write(f"{k}\t{list(v.shape)}\n")
    return base


def variable_to_shape_map(checkpoint_path: str, epoch: int) -> Dict[str, list]:
    """Analogue of NewCheckpointReader(...).get_variable_to_shape_map() (trainers.py:452-454)."""
    out = {}
    with open(f"{checkpoint_path}-{epoch}.index") as f:
        for line in f:
            k, s = line.rstrip("\n").split("\t")
            out[k] = [int(v) for v in re.findall(r"\d+", s)]
    return out


def load_checkpoint(engine, checkpoint_path: str, epoch: int, reuse_vars_scope: Optional[str] = "seq2seq",
                    restore_optimizer: bool = True):
    """Restore the tensors whose name matches reuse_vars_scope (trainers.py:313,348,353,360);
    tensors whose stored shape differs (another subject's conv) are left untouched."""
    with np.load(f"{checkpoint_path}-{epoch}.npz") as z:
        pat = re.compile(reuse_vars_scope) if reuse_vars_scope is not None else None
        restored = []
        for name, (shape, _) in engine.tensors().items():
            if pat is None or not pat.match(name) or name not in z.files or tuple(z[name].shape) != shape:
                continue
            engine.set(name, z[name], L.VALUE)
            if name + EMA_SUFFIX in z.files:
                engine.set(name, z[name + EMA_SUFFIX], L.EMA)
            if restore_optimizer and name + "/Adam" in z.files:
                engine.set(name, z[name + "/Adam"], L.ADAM_M)
                engine.set(name, z[name + "/Adam_1"], L.ADAM_V)
            restored.append(name)
        if restore_optimizer and "global_step" in z.files:
            engine.step = int(z["global_step"])
    return restored

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


def init_engine(engine, seed: int = 1, keep=()):
    """Fresh variables (Glorot / zero biases; the EMA shadow follows the value), zeroed Adam slots.  Tensors named in `keep`
    (just restored from a checkpoint) are left alone.  The reference builds and initialises a new graph for every fit; a
    cached engine must not carry weights or optimiser state over from the previous one."""
    shapes = {k: s for k, (s, _) in engine.tensors().items()}
    fresh = glorot_init(shapes, seed)
    keep = set(keep)
    for name, val in fresh.items():
        if name in keep:
            continue
        engine.set(name, val, L.VALUE)
        z = np.zeros_like(val)
        engine.set(name, z, L.ADAM_M)
        engine.set(name, z, L.ADAM_V)


EMA_SUFFIX = "/ExponentialMovingAverage"  # trainers.py:466-468


def checkpoint_epochs(checkpoint_path: str):
    """Sorted epochs of the checkpoints <checkpoint_path>-<epoch>.index on disk (trainers.py:240-252)."""
    d, base = os.path.split(os.path.abspath(checkpoint_path))
    out = []
    if os.path.isdir(d):
        for f in os.listdir(d):
            m = re.fullmatch(re.escape(base) + r"-(\d+)\.index", f)
            if m:
                out.append(int(m.group(1)))
    return sorted(out)


def save_checkpoint(engine, checkpoint_path: str, epoch: int, max_to_keep: Optional[int] = 5) -> str:
    """<checkpoint_path>-<epoch>.index (manifest, discovered by restore_epoch: trainers.py:240-252)
    + <checkpoint_path>-<epoch>.npz holding variables, EMA shadows and Adam slots under TF names.
    Both files are written to a temporary name and renamed (the .index last: a checkpoint exists once its index does), and
    only the `max_to_keep` most recent checkpoints are kept, like the tf.train.Saver this replaces (None keeps all)."""
    arrays = {}
    for name in engine.tensors():
        arrays[name] = engine.get(name, L.VALUE)
        arrays[name + EMA_SUFFIX] = engine.get(name, L.EMA)
        arrays[name + "/Adam"] = engine.get(name, L.ADAM_M)
        arrays[name + "/Adam_1"] = engine.get(name, L.ADAM_V)
    arrays["global_step"] = np.asarray(engine.step, np.int64)
    base = f"{checkpoint_path}-{epoch}"
    os.makedirs(os.path.dirname(os.path.abspath(base)), exist_ok=True)
    tmp = base + ".tmp.npz"
    np.savez(tmp, **arrays)
    os.replace(tmp, base + ".npz")
    with open(base + ".index.tmp", "w") as f:
        for k, v in arrays.items():
            f.write(f"{k}\t{list(v.shape)}\n")
    os.replace(base + ".index.tmp", base + ".index")
    if max_to_keep:
        for old in checkpoint_epochs(checkpoint_path)[:-int(max_to_keep)]:
            for ext in (".index", ".npz"):      # index first: a half-deleted checkpoint is never discoverable
                try:
                    os.remove(f"{checkpoint_path}-{old}{ext}")
                except OSError:
                    pass
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

import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import parity_common as pc
from ecog2txt_b200 import _lib
from oracle import seq2seq_oracle as O
geo = pc.WIDE
B, T, L = 40, 100, 5
ocfg = O.OracleConfig(**geo)
P = pc.make_params(ocfg)
x, lens, y = pc.make_batch(ocfg, B, T, L, seed=3)
out = {}
for backend in ("simt", "auto"):
    eng = pc.engine_for(geo, _lib.load(), B, T, L, gemm_backend=backend)
    eng.set_all({k: v.numpy() for k, v in P.items()})
    eng.train_step_grads(x, None, y, seed=1)
    out[backend] = eng.get_all(_lib.GRAD)
    eng.close()
k = [n for n in out["simt"] if "encoder_embedding" in n and n.endswith("weights")][0]
a, b = out["simt"][k][0], out["auto"][k][0]     # [W, C, E]
d = np.abs(a - b)
print(k, a.shape, "max", d.max(), "ref max", np.abs(a).max())
print("err by w:", d.max(axis=(1, 2)))
print("err by c:", d.max(axis=(0, 2)))
print("err by e (first 40):", d.max(axis=(0, 1))[:40])
print("err by e (last 40):", d.max(axis=(0, 1))[-40:])
np.set_printoptions(linewidth=200, precision=5, suppress=True)
print("err by e all:", d.max(axis=(0, 1)))
print("ref max by e:", np.abs(a).max(axis=(0, 1)))
e = int(d.max(axis=(0, 1)).argmax())
print("worst e", e, "err by w,c block:\n", d[:, :, e])

#!/bin/bash
mkdir -p gpurun_out
cd tests
cat > /tmp/bis.py <<PY
import sys; sys.path.insert(0,".."); sys.path.insert(0,".")
import numpy as np
import parity_common as pc
from ecog2txt_b200 import _lib, Engine, EngineConfig
from oracle import seq2seq_oracle as O
lib=_lib.load()
def run(name, geo, B, T, L, ragged, train):
    ocfg=O.OracleConfig(**geo)
    x,lens,y=pc.make_batch(ocfg,B,T,L,ragged=ragged)
    eng=pc.engine_for(geo, lib, B, T, L, gemm_backend="auto")
    P=pc.make_params(ocfg); eng.set_all({k:v.numpy() for k,v in P.items()})
    try:
        if train: print(name, eng.train_step_grads(x,None,y,seed=1), flush=True)
        else: print(name, eng.eval_loss(x,None,y), flush=True)
    except Exception as e:
        print(name, "FAIL", repr(e)[:200], flush=True)
        raise SystemExit(0)
    eng.close()
w=sys.argv[1]
if w=="a": run("FULL B256 T400 full eval", pc.FULL, 256, 400, 11, False, False)
if w=="b": run("WIDE B128 T100 ragged train", pc.WIDE, 128, 100, 5, True, True)
if w=="c": run("FULL B256 T400 ragged train", pc.FULL, 256, 400, 11, True, True)
PY
for w in a b c; do E2T_REC_TRAPINFO=1 timeout 120 python /tmp/bis.py $w 2>&1 | tail -3 | cut -c1-300; done
echo "--- without trapinfo"
for w in a b; do timeout 120 python /tmp/bis.py $w 2>&1 | tail -2 | cut -c1-300; done

"""-m gpu: the CUDA path (through the C-ABI) against the oracle on a real B200."""
import numpy as np
import pytest

import parity_common as pc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("backend", ["simt", "auto"])
@pytest.mark.parametrize("geo,B,T,L", [(pc.TINY, 3, 19, 5), (pc.SMALL, 8, 50, 6), (pc.MEDIUM, 32, 120, 7)])
def test_train_step_matches_oracle(gpu_lib, backend, geo, B, T, L):
    # fp32 SIMT: 2e-4 relative; tcgen05 kind::tf32 (10-bit mantissa operands, fp32 accumulate): 1e-2
    tol = 2e-4 if backend == "simt" else 1e-2
    pc.check_train_step(gpu_lib, geo, B, T, L, backend=backend, tol=tol)


@pytest.mark.parametrize("backend", ["simt", "auto"])
def test_train_step_with_dropout(gpu_lib, backend):
    tol = 2e-4 if backend == "simt" else 1e-2
    pc.check_train_step(gpu_lib, pc.MEDIUM, 16, 96, 6, ff=0.1, rnn=0.5, backend=backend, tol=tol)


def test_train_step_explicit_lengths_and_second_subject(gpu_lib):
    pc.check_train_step(gpu_lib, pc.TWO_SUBJ, 4, 17, 5, give_lens=True, subnet=1)


@pytest.mark.parametrize("backend", ["simt", "auto"])
def test_greedy_decode(gpu_lib, backend):
    pc.check_decode(gpu_lib, pc.MEDIUM, 24, 100, 8, backend=backend)


@pytest.mark.parametrize("backend", ["simt", "auto"])
def test_beam_decode(gpu_lib, backend):
    pc.check_decode(gpu_lib, pc.MEDIUM, 6, 60, 6, beam=4, backend=backend)


@pytest.mark.parametrize("B,T,ff,rnn", [(40, 100, 0.0, 0.0), (160, 60, 0.1, 0.5), (129, 30, 0.0, 0.0)])
def test_persistent_recurrent_kernels_full_width(gpu_lib, B, T, ff, rnn):
    """H=400 BiLSTM layers through the whole-sequence tcgen05 kernels (1-2 batch tiles, ragged lengths,
    dropout copies), against the oracle; tolerance = tf32 operands (10-bit mantissa), fp32 accumulate."""
    pc.check_train_step(gpu_lib, pc.WIDE, B, T, 5, ff=ff, rnn=rnn, backend="auto", tol=1e-2)
    c = pc.check_train_step.last_counters
    assert c["persistent_rnn_launches"] == 4, c   # 2 layers x (forward + backward)


def test_persistent_kernels_are_deterministic(gpu_lib):
    import numpy as np
    from ecog2txt_b200 import _lib
    from oracle import seq2seq_oracle as O
    ocfg = O.OracleConfig(**pc.WIDE)
    P = pc.make_params(ocfg)
    x, lens, y = pc.make_batch(ocfg, 64, 80, 5)
    outs = []
    for _ in range(2):
        eng = pc.engine_for(pc.WIDE, gpu_lib, 64, 80, 5, gemm_backend="auto")
        eng.set_all({k: v.numpy() for k, v in P.items()})
        eng.train_step_grads(x, None, y, seed=1)
        outs.append(eng.get_all(_lib.GRAD))
        eng.close()
    for k in outs[0]:
        if "decoder_embedding" in k and k.endswith("weights"):
            continue   # atomicAdd scatter: order-dependent rounding
        assert np.array_equal(outs[0][k], outs[1][k]), k

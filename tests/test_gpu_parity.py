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

"""Engine: object wrapper over the C-ABI handle (include/e2t.h).  No arithmetic happens here."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L


@dataclass
class EngineConfig:
    """Geometry of the network; defaults are subject 400 of the reference's manifest
    (/root/reference/ecog2txt/auxiliary/EFC/mochastar_word_sequence.yaml:62-75,84-85,89)."""
    subnet_ids: Sequence[int] = (400,)
    subnet_C: Sequence[int] = (256,)
    subnet_W: Sequence[int] = (12,)
    E: int = 100
    H: Sequence[int] = (400, 400, 400)
    D: int = 150
    Hd: int = 800
    V: int = 1806
    conv_act: str = "relu"
    emb_act: str = "relu"
    pad_id: int = 0
    eos_id: int = 1
    start_id: int = 1
    max_B: int = 256
    max_T: int = 400
    max_L: int = 20
    max_beam: int = 8
    ff_dropout: float = 0.0
    rnn_dropout: float = 0.0
    lr: float = 5e-4
    beta1: float = 0.9
    beta2: float = 0.999
    eps: float = 1e-8
    ema_decay: float = 0.99
    penalty_scale: float = 1.0
    gemm_backend: str = "auto"
    device: int = 0
    attention: str = "none"    # "luong": optional A7 module (not in the reference model)
    # A6 encoder-targets head ('encoder_1_targets', /root/reference/ecog2txt/trainers.py:798-799; yaml:54,68-69,81); aux_F = 0: none
    aux_layer: int = 1
    aux_hidden: int = 0
    aux_F: int = 0
    aux_kind: str = "gaussian"
    aux_penalty: float = 1.0
    # layer_sizes['decoder_projection'] (yaml:65, empty in the shipped manifests): one optional hidden FF layer; 0 = none
    proj_hidden: int = 0

    def to_c(self) -> L.E2TConfig:
        c = L.E2TConfig()
        n = len(self.subnet_ids)
        if not (len(self.subnet_C) == len(self.subnet_W) == n):
            raise ValueError("subnet_ids / subnet_C / subnet_W must have equal lengths")
        if n > L.E2T_MAX_SUBNETS or len(self.H) > L.E2T_MAX_LAYERS:
            raise ValueError("too many subnets / encoder layers")
        c.n_subnets = n
        for i in range(n):
            c.subnet_id[i], c.subnet_C[i], c.subnet_W[i] = int(self.subnet_ids[i]), int(self.subnet_C[i]), int(self.subnet_W[i])
        c.E, c.n_enc_layers = int(self.E), len(self.H)
        for i, hh in enumerate(self.H):
            c.H[i] = int(hh)
        c.D, c.Hd, c.V = int(self.D), int(self.Hd), int(self.V)
        c.conv_act, c.emb_act = L.ACT[self.conv_act], L.ACT[self.emb_act]
        c.pad_id, c.eos_id, c.start_id = self.pad_id, self.eos_id, self.start_id
        c.max_B, c.max_T, c.max_L, c.max_beam = self.max_B, self.max_T, self.max_L, self.max_beam
        c.ff_dropout, c.rnn_dropout = self.ff_dropout, self.rnn_dropout
        c.lr, c.beta1, c.beta2, c.eps = self.lr, self.beta1, self.beta2, self.eps
        c.ema_decay, c.penalty_scale = self.ema_decay, self.penalty_scale
        c.gemm_backend, c.device = L.GEMM[self.gemm_backend], self.device
        c.attention = L.ATTN[self.attention]
        c.aux_layer, c.aux_hidden, c.aux_F = int(self.aux_layer), int(self.aux_hidden), int(self.aux_F)
        c.aux_kind, c.aux_penalty = L.AUX_KIND[self.aux_kind], float(self.aux_penalty)
        c.proj_hidden = int(self.proj_hidden)
        return c


class E2TError(RuntimeError):
    pass


def _ptr(a):
    """(void*, loc) of a numpy array (host) or of anything with data_ptr() on a CUDA device."""
    if a is None:
        return None, None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p), L.HOST
    if hasattr(a, "data_ptr"):
        if not a.is_cuda:
            raise TypeError("torch inputs must live on the GPU; pass numpy arrays for host data")
        if not a.is_contiguous():
            raise TypeError("device inputs must be contiguous")
        return C.c_void_p(a.data_ptr()), L.DEVICE
    raise TypeError(f"unsupported buffer type {type(a)}")


class Engine:
    """One handle <-> one GPU <-> one stream.  `lib` is injectable for the kernel-emulation tests only."""

    def __init__(self, cfg: EngineConfig, lib: Optional[C.CDLL] = None):
        self.cfg = cfg
        self._lib = lib if lib is not None else L.load()
        self.emulated = lib is not None
        self._h = C.c_void_p()
        cc = cfg.to_c()
        if self._lib.e2t_create(C.byref(cc), C.byref(self._h)) != 0:
            raise E2TError(self._lib.e2t_last_error().decode())
        self._names: Optional[Dict[str, Tuple[Tuple[int, ...], int]]] = None

    # -- plumbing ---------------------------------------------------------------------------
    def _ck(self, rc):
        if rc != 0:
            raise E2TError(self._lib.e2t_last_error().decode())

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._lib.e2t_sync(self._h)
            for p in getattr(self, "_host_bufs", []):
                self._lib.e2t_host_free(p)
            self._host_bufs = []
            self._lib.e2t_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream: int):
        self._ck(self._lib.e2t_set_stream(self._h, C.c_void_p(cuda_stream)))

    def sync(self):
        self._ck(self._lib.e2t_sync(self._h))

    # -- parameters -------------------------------------------------------------------------
    def tensors(self) -> Dict[str, Tuple[Tuple[int, ...], int]]:
        """{name: (shape, offset)} in flat-buffer order."""
        if self._names is None:
            out = {}
            n = self._lib.e2t_param_count(self._h)
            buf = C.create_string_buffer(256)
            shape = (C.c_int64 * 4)()
            nd, off = C.c_int(), C.c_int64()
            for i in range(n):
                self._ck(self._lib.e2t_param_info(self._h, i, buf, 256, shape, C.byref(nd), C.byref(off)))
                out[buf.value.decode()] = (tuple(int(shape[j]) for j in range(nd.value)), int(off.value))
            self._names = out
        return self._names

    @property
    def n_params(self) -> int:
        return sum(int(np.prod(s)) for s, _ in self.tensors().values())

    def get(self, name: str, which: int = L.VALUE) -> np.ndarray:
        shape, _ = self.tensors()[name]
        out = np.empty(shape, np.float32)
        self._ck(self._lib.e2t_get_tensor(self._h, name.encode(), which, out.ctypes.data_as(C.c_void_p)))
        return out

    def set(self, name: str, value, which: int = L.VALUE):
        shape, _ = self.tensors()[name]
        a = np.ascontiguousarray(np.asarray(value, np.float32))
        if a.shape != shape:
            raise ValueError(f"{name}: shape {a.shape} != {shape}")
        self._ck(self._lib.e2t_set_tensor(self._h, name.encode(), which, a.ctypes.data_as(C.c_void_p)))

    def set_all(self, params: Dict[str, np.ndarray], which: int = L.VALUE):
        for k in self.tensors():
            self.set(k, params[k], which)

    def get_all(self, which: int = L.VALUE) -> Dict[str, np.ndarray]:
        return {k: self.get(k, which) for k in self.tensors()}

    def flat_buffer(self, which: int = L.GRAD) -> Tuple[int, int]:
        p, n = C.c_void_p(), C.c_int64()
        self._ck(self._lib.e2t_flat_buffer(self._h, which, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def set_trainable(self, name: str, flag: bool):
        self._ck(self._lib.e2t_set_trainable(self._h, name.encode(), int(bool(flag))))

    @property
    def step(self) -> int:
        s = C.c_int64()
        self._ck(self._lib.e2t_get_step(self._h, C.byref(s)))
        return int(s.value)

    @step.setter
    def step(self, v: int):
        self._ck(self._lib.e2t_set_step(self._h, int(v)))

    # -- hot path ---------------------------------------------------------------------------
    def _inputs(self, x, lens, y=None):
        px, loc = _ptr(x)
        B, T = int(x.shape[0]), int(x.shape[1])
        pl, locl = _ptr(lens)
        py, locy = _ptr(y)
        for l2 in (locl, locy):
            if l2 is not None and l2 != loc:
                raise TypeError("x, lens and y must all be host arrays or all device tensors")
        if isinstance(x, np.ndarray):
            if x.dtype != np.float32 or not x.flags.c_contiguous:
                raise TypeError("x must be C-contiguous float32")
            for a in (lens, y):
                if a is not None and (a.dtype != np.int32 or not a.flags.c_contiguous):
                    raise TypeError("lens / y must be C-contiguous int32")
        return px, pl, py, loc, B, T

    def train_step_grads(self, x, lens, y, subnet: int = 0, seed: int = 0, want_loss: bool = True):
        """forward + backward; returns (loss_sum, ntok) or None (no host sync) if not want_loss."""
        px, pl, py, loc, B, T = self._inputs(x, lens, y)
        Lk = int(y.shape[1])
        if want_loss:
            loss, ntok = C.c_float(), C.c_int32()
            self._ck(self._lib.e2t_train_step_grads(self._h, subnet, px, pl, py, loc, B, T, Lk, seed & 0xFFFFFFFF,
                                                    C.byref(loss), C.byref(ntok)))
            return float(loss.value), int(ntok.value)
        self._ck(self._lib.e2t_train_step_grads(self._h, subnet, px, pl, py, loc, B, T, Lk, seed & 0xFFFFFFFF, None, None))
        return None

    def stage_inputs(self, slot: int, x: np.ndarray, lens, y, subnet: int = 0):
        """Start the host->device copy of a minibatch into staging slot 0/1 on the library's copy stream (returns at once
        when the arrays are page-locked); consume it with train_step_grads_staged(slot, ...)."""
        if not isinstance(x, np.ndarray):
            raise TypeError("stage_inputs takes host (numpy) arrays")
        px, pl, py, _, B, T = self._inputs(x, lens, y)
        Lk = int(y.shape[1]) if y is not None else 0
        self._ck(self._lib.e2t_stage_inputs(self._h, slot, subnet, px, pl, py, B, T, Lk))
        self._staged_shape = getattr(self, "_staged_shape", {})
        self._staged_shape[slot] = (B, T, Lk, subnet)

    def train_step_grads_staged(self, slot: int, seed: int = 0, want_loss: bool = True):
        B, T, Lk, subnet = self._staged_shape[slot]
        loc = L.STAGED0 + slot
        if want_loss:
            loss, ntok = C.c_float(), C.c_int32()
            self._ck(self._lib.e2t_train_step_grads(self._h, subnet, None, None, None, loc, B, T, Lk, seed & 0xFFFFFFFF,
                                                    C.byref(loss), C.byref(ntok)))
            return float(loss.value), int(ntok.value)
        self._ck(self._lib.e2t_train_step_grads(self._h, subnet, None, None, None, loc, B, T, Lk, seed & 0xFFFFFFFF, None, None))
        return None

    def host_buffer(self, shape, dtype=np.float32) -> np.ndarray:
        """numpy array over page-locked host memory (e2t_host_alloc), freed when the engine is closed: a staging buffer from
        which stage_inputs copies asynchronously."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self._ck(self._lib.e2t_host_alloc(C.byref(p), n))
        self._host_bufs = getattr(self, "_host_bufs", [])
        self._host_bufs.append(p)
        buf = (C.c_char * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def set_grad_buckets(self, on: bool):
        """complete the gradient buffer bucket by bucket (with an event each) so that the all-reduce can overlap the backward"""
        self._ck(self._lib.e2t_set_grad_buckets(self._h, int(bool(on))))

    def grad_buckets(self) -> List[Tuple[int, int]]:
        """[(offset, n)] flat ranges of E2T_GRAD in the order the most recent train_step_grads completes them"""
        out = []
        off, n = C.c_int64(), C.c_int64()
        for i in range(self._lib.e2t_grad_bucket_count(self._h)):
            self._ck(self._lib.e2t_grad_bucket_info(self._h, i, C.byref(off), C.byref(n)))
            out.append((int(off.value), int(n.value)))
        return out

    def grad_bucket_wait(self, i: int, cuda_stream: int):
        self._ck(self._lib.e2t_grad_bucket_wait(self._h, i, C.c_void_p(cuda_stream)))

    def set_encoder_targets(self, targets):
        """A6: encoder targets of the next train_step_grads / eval_loss / input_saliency call: float32 [B,T,aux_F]
        (gaussian) or int32 [B,T] (categorical), host (numpy) or device (torch.cuda)."""
        p, loc = _ptr(targets)
        want = np.float32 if self.cfg.aux_kind == "gaussian" else np.int32
        if isinstance(targets, np.ndarray) and (targets.dtype != want or not targets.flags.c_contiguous):
            raise TypeError(f"encoder targets must be C-contiguous {np.dtype(want).name}")
        self._ck(self._lib.e2t_set_encoder_targets(self._h, p, loc, int(targets.shape[0]), int(targets.shape[1])))

    def last_losses(self):
        """(decoder loss sum, ntok, encoder-targets loss sum, unmasked frames) of the most recent step."""
        ld, la, nt, nf = C.c_float(), C.c_float(), C.c_int32(), C.c_int32()
        self._ck(self._lib.e2t_last_losses(self._h, C.byref(ld), C.byref(nt), C.byref(la), C.byref(nf)))
        return float(ld.value), int(nt.value), float(la.value), int(nf.value)

    def post_losses(self, slot: int):
        """Enqueue the copy of the most recent step's losses into page-locked ring slot 0..3 (no synchronisation)."""
        self._ck(self._lib.e2t_post_losses(self._h, int(slot)))

    def fetch_losses(self, slot: int):
        """Wait for the copy posted into `slot` and return (decoder loss sum, ntok, encoder-targets loss sum, frames)."""
        ld, la, nt, nf = C.c_float(), C.c_float(), C.c_int32(), C.c_int32()
        self._ck(self._lib.e2t_fetch_losses(self._h, int(slot), C.byref(ld), C.byref(nt), C.byref(la), C.byref(nf)))
        return float(ld.value), int(nt.value), float(la.value), int(nf.value)

    def input_saliency(self, x, lens, y, subnet: int = 0, use_ema: bool = False, decoder_penalty: Optional[float] = None,
                       aux_penalty: Optional[float] = None, want_dx: bool = True, want_norms: bool = True):
        """A13: d(loss)/d(encoder_inputs) [B,T,C] and its per-electrode squared norms over time [B,C] (host inputs only)."""
        if not isinstance(x, np.ndarray):
            raise TypeError("input_saliency takes host (numpy) arrays")
        px, pl, py, loc, B, T = self._inputs(x, lens, y)
        Cc = int(x.shape[2])
        dx = np.empty((B, T, Cc), np.float32) if want_dx else None
        sq = np.empty((B, Cc), np.float32) if want_norms else None
        pd = self.cfg.penalty_scale if decoder_penalty is None else decoder_penalty
        pa = self.cfg.aux_penalty if aux_penalty is None else aux_penalty
        self._ck(self._lib.e2t_input_saliency(self._h, subnet, px, pl, py, loc, B, T, int(y.shape[1]), int(use_ema), pd, pa,
                                              dx.ctypes.data_as(C.c_void_p) if want_dx else None,
                                              sq.ctypes.data_as(C.c_void_p) if want_norms else None))
        return dx, sq

    def adam_ema_step(self, grad_scale: float, subnet: int = -1):
        self._ck(self._lib.e2t_adam_ema_step(self._h, subnet, float(grad_scale)))

    def adam_ema_step_dev(self, token_count=None, subnet: int = -1):
        """Adam + EMA with grad_scale = 1 / max(token_count, 1) read on the device: `token_count` is a 1-element float32
        device tensor (numpy in the emulation build), e.g. the all-reduced token count -- no host synchronisation.
        None: the count slot behind the gradient buffer (flat_buffer(GRAD_AND_COUNT), all-reduced with the gradients)."""
        p = None if token_count is None else _ptr(token_count)[0]
        self._ck(self._lib.e2t_adam_ema_step_dev(self._h, subnet, p))

    def read_loss_accumulators(self, reset: bool = True):
        """(decoder loss sum, tokens, encoder-targets loss sum, frames) over the training steps since the last reset;
        one host synchronisation."""
        out = (C.c_double * 4)()
        self._ck(self._lib.e2t_read_loss_accumulators(self._h, out, int(reset)))
        return float(out[0]), int(round(out[1])), float(out[2]), int(round(out[3]))

    def wait_staged(self, slot: int):
        """Block until the latest stage_inputs copy into `slot` has finished (its host buffer may then be rewritten)."""
        self._ck(self._lib.e2t_wait_staged(self._h, int(slot)))

    def eval_loss(self, x, lens, y, subnet: int = 0, use_ema: bool = False):
        px, pl, py, loc, B, T = self._inputs(x, lens, y)
        loss, ntok = C.c_float(), C.c_int32()
        self._ck(self._lib.e2t_eval_loss(self._h, subnet, px, pl, py, loc, B, T, int(y.shape[1]), int(use_ema),
                                         C.byref(loss), C.byref(ntok)))
        return float(loss.value), int(ntok.value)

    def greedy_decode(self, x, lens=None, max_len: int = 20, subnet: int = 0, use_ema: bool = False,
                      temperature: float = 1.0, want_logp: bool = True):
        px, pl, _, loc, B, T = self._inputs(x, lens)
        toks = np.empty((B, max_len), np.int32)
        logp = np.empty((B, max_len), np.float32) if want_logp else None
        self._ck(self._lib.e2t_greedy_decode(self._h, subnet, px, pl, loc, B, T, max_len, int(use_ema), temperature,
                                             toks.ctypes.data_as(C.c_void_p),
                                             logp.ctypes.data_as(C.c_void_p) if want_logp else None))
        return toks, logp

    def beam_decode(self, x, lens=None, beam: int = 8, max_len: int = 20, subnet: int = 0, use_ema: bool = False,
                    temperature: float = 1.0):
        px, pl, _, loc, B, T = self._inputs(x, lens)
        toks = np.empty((B, beam, max_len), np.int32)
        scores = np.empty((B, beam), np.float32)
        self._ck(self._lib.e2t_beam_decode(self._h, subnet, px, pl, loc, B, T, beam, max_len, int(use_ema), temperature,
                                           toks.ctypes.data_as(C.c_void_p), scores.ctypes.data_as(C.c_void_p)))
        return toks, scores

    def activation(self, name: str, shape, dtype=np.float32) -> np.ndarray:
        out = np.empty(shape, dtype)
        n = C.c_int64()
        self._ck(self._lib.e2t_get_activation(self._h, name.encode(), out.ctypes.data_as(C.c_void_p), out.size, C.byref(n)))
        if n.value != out.size:
            raise E2TError(f"activation {name}: got {n.value} elements, expected {out.size}")
        return out

    def launch_counts(self) -> Tuple[int, int]:
        a, b = C.c_int64(), C.c_int64()
        self._ck(self._lib.e2t_launch_counts(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def counter(self, name: str) -> int:
        v = C.c_int64()
        self._ck(self._lib.e2t_counter(self._h, name.encode(), C.byref(v)))
        return int(v.value)

    def profile_enable(self, on: bool):
        self._ck(self._lib.e2t_profile_enable(self._h, int(on)))

    def profile_read(self, category: int) -> Tuple[float, int]:
        ms, n = C.c_double(), C.c_int64()
        self._ck(self._lib.e2t_profile_read(self._h, category, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def profile_report(self) -> Dict[str, Tuple[int, float]]:
        """{kernel label: (launches, total ms)} of the records taken since profile_enable(True)."""
        buf = C.create_string_buffer(1 << 16)
        self._ck(self._lib.e2t_profile_report(self._h, buf, len(buf)))
        out = {}
        for line in buf.value.decode().splitlines():
            k, n, ms = line.split("\t")
            out[k] = (int(n), float(ms))
        return out

    def bench_gemm(self, M: int, N: int, K: int, tn: bool = False, beta: float = 0.0, iters: int = 20) -> float:
        ms = C.c_float()
        self._ck(self._lib.e2t_bench_gemm(self._h, M, N, K, int(tn), beta, iters, C.byref(ms)))
        return float(ms.value)

    def selftest_gemm(self, M: int, N: int, K: int) -> float:
        d = C.c_float()
        self._ck(self._lib.e2t_selftest_gemm(self._h, M, N, K, C.byref(d)))
        return float(d.value)

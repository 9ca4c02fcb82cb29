"""MiniROAD online step-recognition module backed by the sm_100a C-ABI library.

Drop-in for the reference ``MROAD`` (``step_recognition/model/rnn/rnn.py:18-71``):

* same constructor contract ``MROAD(cfg: dict)`` reading ``no_flow, no_rgb, rgb_type,
  flow_type, hidden_dim, num_layers, num_classes, window_size, embedding_dim, dropout``
  (rnn.py:23-43);
* same ten ``state_dict`` tensors (``gru.*``, ``layer1.0.*``, ``layer1.1.*``,
  ``f_classification.0.*``), created in the reference's order so that the same
  ``torch.manual_seed`` gives the same default initialisation; ``h0`` is a plain attribute;
* same ``forward(rgb_input, flow_input) -> {'logits': Tensor[B, T, K]}`` returning softmax
  probabilities in eval mode (rnn.py:65-71).

The torch sub-modules are parameter containers only -- their ``forward`` is never called.
All arithmetic runs in ``libprego_b200.so``; there is no PyTorch / CPU fallback and CPU
tensors are rejected loudly.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from .registry import META_ARCHITECTURES

# rnn.py:6-16
FEATURE_SIZES = {
    "rgb_anet_resnet50": 2048,
    "flow_anet_resnet50": 2048,
    "rgb_kinetics_bninception": 1024,
    "flow_kinetics_bninception": 1024,
    "rgb_kinetics_resnet50": 2048,
    "flow_kinetics_resnet50": 2048,
    "flow_nv_kinetics_bninception": 1024,
    "rgb_kinetics_i3d": 2048,
    "flow_kinetics_i3d": 2048,
}

_STATE_KEYS = (
    "layer1.0.weight", "layer1.0.bias", "layer1.1.weight", "layer1.1.bias",
    "gru.weight_ih_l0", "gru.weight_hh_l0", "gru.bias_ih_l0", "gru.bias_hh_l0",
    "f_classification.0.weight", "f_classification.0.bias",
)


@META_ARCHITECTURES.register("MiniROAD")
class MROAD(nn.Module):
    """B200-native MiniROAD.  Extra (optional) cfg keys: ``precision`` ('fp16' default: throughput | 'bf16' | 'fp16x3': fp32-class
    accuracy on the tensor cores (fp16 hi + lo operands) | 'fp32': exact CUDA-core FFMA),
    ``chunk_frames`` (frames per pass over all streams; bounds the workspace), ``train_precision`` ('tf32x3' default:
    the large projections and their gradients on tcgen05 kind::tf32 over split hi + lo operands, fp32-class forward, 1.6x the stock
    torch step | 'fp32': exact CUDA-core GEMMs, the gradient-parity mode | 'tf32': plain TF32 operands, the accuracy of stock torch's
    default cuDNN GRU; shapes the tensor-core GEMM does not take (B*T % 32 != 0) run the exact kernels in every mode)."""

    def __init__(self, cfg):
        super().__init__()
        self.use_flow = not cfg["no_flow"]
        self.use_rgb = not cfg["no_rgb"]
        self.d_rgb = FEATURE_SIZES[cfg["rgb_type"]] if self.use_rgb else 0
        self.d_flow = FEATURE_SIZES[cfg["flow_type"]] if self.use_flow else 0
        self.input_dim = self.d_rgb + self.d_flow
        self.hidden_dim = cfg["hidden_dim"]
        self.num_layers = cfg["num_layers"]
        self.out_dim = cfg["num_classes"]
        self.window_size = cfg["window_size"]
        self.embedding_dim = cfg["embedding_dim"]
        if self.num_layers != 1:
            raise ValueError("prego_b200 implements the shipped 1-layer GRU (configs/*.yaml: num_layers 1)")
        # parameter containers, created in the reference's order (rnn.py:38-47)
        self.gru = nn.GRU(self.embedding_dim, self.hidden_dim, self.num_layers, batch_first=True)
        self.layer1 = nn.Sequential(
            nn.Linear(self.input_dim, self.embedding_dim),
            nn.LayerNorm(self.embedding_dim),
            nn.ReLU(),
            nn.Dropout(p=cfg["dropout"]),
        )
        self.f_classification = nn.Sequential(nn.Linear(self.hidden_dim, self.out_dim))
        self.h0 = torch.zeros(self.num_layers, 1, self.hidden_dim)  # rnn.py:49 (not in state_dict)

        self.precision = cfg.get("precision", "fp16")
        self.train_precision = cfg.get("train_precision", "tf32x3")
        self.chunk_frames = int(cfg.get("chunk_frames", 1 << 17))
        self._handle = None
        self._handle_device = None
        self._packed_key = None
        self._packed_formats = 0
        self._workspace = None
        self._ant_workspace = None
        self.last_labels = None  # int32 [B, T] labels of the last forward (fused argmax)

    # ------------------------------------------------------------------ C-ABI plumbing
    def _release(self):
        if self._handle is not None:
            _lib.load().prego_model_destroy(self._handle)
            self._handle = None
            self._packed_key = None
            self._packed_formats = 0

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _ensure_handle(self, device: torch.device):
        lib = _lib.load()
        if self._handle is not None and self._handle_device == device:
            return lib
        self._release()
        dims = _lib.Dims(self.d_rgb, self.d_flow, self.embedding_dim, self.hidden_dim, self.out_dim)
        h = C.c_void_p()
        _lib.check(lib.prego_model_create(C.byref(dims), device.index or 0, C.byref(h)), "prego_model_create")
        self._handle, self._handle_device = h, device
        return lib

    def _param_tensors(self):
        l1, ln, fc, g = self.layer1[0], self.layer1[1], self.f_classification[0], self.gru
        return (l1.weight, l1.bias, ln.weight, ln.bias, g.weight_ih_l0, g.weight_hh_l0, g.bias_ih_l0, g.bias_hh_l0,
                fc.weight, fc.bias)  # order of _STATE_KEYS / prego_weights_t

    def _sync_weights(self, lib, device, formats=_lib.PACK_ALL):
        """Re-pack the ten tensors into the library handle when any of them changed (in-place update,
        load_state_dict, .to()) or when a format this call needs is stale.  Cheap when nothing changed: ten
        (data_ptr, version) pairs.  ``formats``: the PREGO_PACK_* set the caller reads -- the training step passes the
        fp32 set only (it re-packs after every optimizer step); inference packs everything, once."""
        tensors = self._param_tensors()
        key = tuple((t.data_ptr(), t._version) for t in tensors)
        if key == self._packed_key and (self._packed_formats & formats) == formats:
            return
        for k, t in zip(_STATE_KEYS, tensors):
            if t.device != device or t.dtype != torch.float32:
                raise RuntimeError(f"parameter {k} must be fp32 on {device} (got {t.dtype} on {t.device}); call model.to(device)")
        if key == self._packed_key:
            formats |= self._packed_formats  # same weights: what is already packed stays valid (the call marks the rest stale)
        keep = [t.detach().contiguous() for t in tensors]
        w = _lib.Weights(*[t.data_ptr() for t in keep])
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(lib.prego_model_load_weights_ex(self._handle, C.byref(w), formats, stream), "prego_model_load_weights_ex")
        self._packed_key, self._packed_formats = key, formats | _lib.PACK_F32

    def _get_workspace(self, nbytes: int, device):
        ws = self._workspace
        if ws is None or ws.device != device or ws.numel() < nbytes + 1024:
            self._workspace = None
            ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
            self._workspace = ws
        off = (-ws.data_ptr()) % 1024
        return ws.data_ptr() + off

    # ------------------------------------------------------------------------ inference
    @torch.no_grad()
    def infer(self, rgb_input, flow_input, h_state=None, want_probs=True, want_logits=False, want_labels=True,
              precision=None, chunk_T=None, out=None, zero_flow=False, want_anticipation=False,
              want_anticipation_logits=False):
        """Run the CUDA path.  rgb/flow: CUDA tensors [B, T, D], fp32 (the reference loader's format) or already in the
        16-bit operand format of ``precision`` (torch.float16 for 'fp16', torch.bfloat16 for 'bf16': same results, no
        staging pass, half the bytes).  ``zero_flow``: the caller asserts the flow stream is all zero, as the shipped
        configs feed it (datasets/dataset.py:63-69); ``flow_input`` may then be None and its half of the projection
        is skipped (bit-identical to passing zeros).  ``h_state`` ([B, H] fp32 CUDA) is updated in place when given
        (streaming / time-chunked online inference).  ``want_anticipation`` / ``want_anticipation_logits`` (MROADA only)
        add ``anticipation_probs`` / ``anticipation_logits`` [B, T, A, K] and ``anticipation_labels`` [B, T, A]."""
        ref = rgb_input if self.use_rgb else flow_input
        if not isinstance(ref, torch.Tensor) or not ref.is_cuda:
            raise RuntimeError("prego_b200.MROAD runs on CUDA (sm_100a) tensors only; there is no CPU fallback")
        device = ref.device
        B, T = int(ref.shape[0]), int(ref.shape[1])
        prec_name = precision or self.precision
        feat_dtype = ref.dtype
        want16 = {"fp16": torch.float16, "bf16": torch.bfloat16}.get(prec_name)
        if feat_dtype != torch.float32 and feat_dtype != want16:
            raise RuntimeError(f"features must be fp32, or {want16} for precision '{prec_name}'; got {feat_dtype}")

        def prep(x, d, name):
            if d == 0:
                return None
            if x.device != device or x.dtype != feat_dtype or tuple(x.shape) != (B, T, d):
                raise RuntimeError(f"{name} must be {feat_dtype} [{B}, {T}, {d}] on {device}, got {x.dtype} {tuple(x.shape)} on {x.device}")
            return x.contiguous()

        rgb = prep(rgb_input, self.d_rgb, "rgb_input")
        flow = None if (zero_flow and flow_input is None) else prep(flow_input, self.d_flow, "flow_input")
        prec = _lib.PRECISIONS[precision or self.precision]
        with torch.cuda.device(device):
            lib = self._ensure_handle(device)
            self._sync_weights(lib, device)
            if chunk_T is None:
                chunk_T = max(1, min(T, self.chunk_frames // max(B, 1)))
            chunk_T = int(min(chunk_T, T))
            need = lib.prego_workspace_bytes(self._handle, B, chunk_T, prec)
            ws_ptr = self._get_workspace(need, device)
            K = self.out_dim
            if out is not None:  # caller-owned output buffers (streaming loops: no allocation per call)
                probs, logits, labels = out.get("probs"), out.get("logits"), out.get("labels")
            else:
                probs = torch.empty(B, T, K, dtype=torch.float32, device=device) if want_probs else None
                logits = torch.empty(B, T, K, dtype=torch.float32, device=device) if want_logits else None
                labels = torch.empty(B, T, dtype=torch.int32, device=device) if want_labels else None
            if h_state is not None:
                if h_state.device != device or h_state.dtype != torch.float32 or tuple(h_state.shape) != (B, self.hidden_dim) \
                        or not h_state.is_contiguous():
                    raise RuntimeError(f"h_state must be contiguous fp32 [{B}, {self.hidden_dim}] on {device}")
            args = _lib.ForwardArgs(
                rgb.data_ptr() if rgb is not None else None, flow.data_ptr() if flow is not None else None, B, T,
                h_state.data_ptr() if h_state is not None else None,
                probs.data_ptr() if probs is not None else None,
                logits.data_ptr() if logits is not None else None,
                labels.data_ptr() if labels is not None else None,
                ws_ptr, need, prec, chunk_T, _lib.FEAT_F32 if feat_dtype == torch.float32 else _lib.FEAT_16, 1 if zero_flow else 0)
            stream = torch.cuda.current_stream(device).cuda_stream
            if not (want_anticipation or want_anticipation_logits):
                _lib.check(lib.prego_forward(self._handle, C.byref(args), stream), "prego_forward")
                return {"probs": probs, "logits": logits, "labels": labels}
            A = int(getattr(self, "anticipation_length", 0))
            if A <= 0:
                raise RuntimeError("anticipation outputs need the MROADA module (model: MiniROADA)")
            ant_probs = torch.empty(B, T, A, K, dtype=torch.float32, device=device) if want_anticipation else None
            ant_logits = torch.empty(B, T, A, K, dtype=torch.float32, device=device) if want_anticipation_logits else None
            ant_labels = torch.empty(B, T, A, dtype=torch.int32, device=device)
            slab = min(B * chunk_T, int(self.anticipation_slab_rows))  # rows of the [rows, A * H] activation alive at once
            ant_need = lib.prego_anticipation_workspace_bytes(self._handle, slab, prec)
            if self._ant_workspace is None or self._ant_workspace.device != device or self._ant_workspace.numel() < ant_need + 1024:
                self._ant_workspace = None
                self._ant_workspace = torch.empty(ant_need + 1024, dtype=torch.uint8, device=device)
            aws = self._ant_workspace.data_ptr() + (-self._ant_workspace.data_ptr()) % 1024
            ant = _lib.AnticipationArgs(ant_probs.data_ptr() if ant_probs is not None else None,
                                        ant_logits.data_ptr() if ant_logits is not None else None,
                                        ant_labels.data_ptr(), aws, ant_need)
            _lib.check(lib.prego_forward_anticipation(self._handle, C.byref(args), C.byref(ant), stream), "prego_forward_anticipation")
        return {"probs": probs, "logits": logits, "labels": labels, "anticipation_probs": ant_probs,
                "anticipation_logits": ant_logits, "anticipation_labels": ant_labels}

    def online_session(self, num_streams: int, device=None, precision=None, want_probs=False, host_labels=False):
        """Strict per-frame online inference: ``session.step(rgb_frame, flow_frame) -> labels`` with the GRU state
        carried inside the session and all per-call host work hoisted out (BASELINE configs[1])."""
        return OnlineSession(self, num_streams, device, precision, want_probs, host_labels)

    def device_error(self) -> int:
        """Watchdog flag of the persistent recurrence kernels (0 = healthy); synchronises the device."""
        if self._handle is None:
            return 0
        v = C.c_int32(0)
        _lib.check(_lib.load().prego_device_error(self._handle, C.byref(v)), "prego_device_error")
        return int(v.value)

    def recurrence_fallbacks(self) -> int:
        """Launches of the batched recurrence that ran without the cooperative attribute (0 on a healthy stack)."""
        if self._handle is None:
            return 0
        v = C.c_int64(0)
        _lib.check(_lib.load().prego_recurrence_fallbacks(self._handle, C.byref(v)), "prego_recurrence_fallbacks")
        return int(v.value)

    def check_device(self, where: str = "") -> None:
        """Raise if a persistent kernel's watchdog fired since the last check: its outputs (labels, mAP, gradients)
        would be garbage.  Called at the end of every product-level loop (evaluation, epoch, batched prediction);
        synchronises the device, so never per step."""
        err = self.device_error()
        if err != 0:
            raise RuntimeError(f"prego_b200: a persistent recurrence kernel timed out waiting for a peer CTA (flag {err})"
                               f"{' during ' + where if where else ''}; results of this run are invalid")

    def profile_begin(self):
        """Arm per-phase CUDA-event timing inside the library (used by bench.py's roofline)."""
        if self._handle is None:
            raise RuntimeError("run one inference first (the handle is created lazily)")
        _lib.check(_lib.load().prego_profile_begin(self._handle), "prego_profile_begin")

    def profile_end(self):
        """-> {phase: {"ms": device time, "launches": kernel launches}} since profile_begin()."""
        n = len(_lib.PHASES)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        _lib.check(_lib.load().prego_profile_end(self._handle, ms, cnt), "prego_profile_end")
        return {p: {"ms": ms[i], "launches": int(cnt[i])} for i, p in enumerate(_lib.PHASES)}

    def forward(self, rgb_input, flow_input):
        """rnn.py:51-71.  Eval: ``out['logits']`` = softmax probabilities [B, T, K]; train: raw logits."""
        if self.training:
            # train mode: raw logits, dropout active, autograd-tracked (rnn.py:66-67, trainer/train.py:20-24)
            from .training import train_forward
            return {"logits": train_forward(self, rgb_input, flow_input)}
        out = self.infer(rgb_input, flow_input, want_probs=True, want_labels=True)
        self.last_labels = out["labels"]
        return {"logits": out["probs"]}


_ANT_KEYS = ("anticipation_layer.0.weight", "anticipation_layer.0.bias")


@META_ARCHITECTURES.register("MiniROADA")
class MROADA(MROAD):
    """B200-native MiniROADA (``rnn.py:73-137``): the MROAD trunk plus the anticipation head
    ``anticipation_layer = Linear(H, A * H)`` whose A hidden vectors per frame go through the SAME classifier.

    Constructor contract of the reference (``rnn.py:76-110``): reads ``no_flow, no_rgb, rgb_type, flow_type,
    embedding_dim, hidden_dim, num_layers, anticipation_length, num_classes, dropout, actionness``; parameters are
    created in the reference's order (layer1, [f_actionness], gru, f_classification, anticipation_layer), so the
    same ``torch.manual_seed`` gives the same default initialisation and ``state_dict`` has the same 12 (14 with
    ``actionness``) tensors.  ``f_actionness`` is a parameter container only, exactly as in the reference, whose
    forward never calls it (``rnn.py:112-137``).  Eval-mode forward returns ``{'logits': softmax [B, T, K],
    'anticipation_logits': softmax [B, T, A, K]}`` (``rnn.py:132-135``).  Inference only: the train-mode forward
    raises (the anticipation training loop and its dataset are outside the hot path)."""

    def __init__(self, cfg):
        nn.Module.__init__(self)
        self.use_flow = not cfg["no_flow"]
        self.use_rgb = not cfg["no_rgb"]
        self.d_rgb = FEATURE_SIZES[cfg["rgb_type"]] if self.use_rgb else 0
        self.d_flow = FEATURE_SIZES[cfg["flow_type"]] if self.use_flow else 0
        self.input_dim = self.d_rgb + self.d_flow
        self.embedding_dim = cfg["embedding_dim"]
        self.hidden_dim = cfg["hidden_dim"]
        self.num_layers = cfg["num_layers"]
        self.anticipation_length = int(cfg["anticipation_length"])
        self.out_dim = cfg["num_classes"]
        if self.num_layers != 1:
            raise ValueError("prego_b200 implements the shipped 1-layer GRU (configs/*.yaml: num_layers 1)")
        if self.anticipation_length < 1:
            raise ValueError("anticipation_length must be >= 1")
        self.layer1 = nn.Sequential(
            nn.Linear(self.input_dim, self.embedding_dim),
            nn.LayerNorm(self.embedding_dim),
            nn.ReLU(),
            nn.Dropout(p=cfg["dropout"]),
        )
        self.actionness = cfg["actionness"]
        if self.actionness:
            self.f_actionness = nn.Sequential(nn.Linear(self.hidden_dim, 1))
        self.relu = nn.ReLU()
        self.gru = nn.GRU(self.embedding_dim, self.hidden_dim, self.num_layers, batch_first=True)
        self.f_classification = nn.Sequential(nn.Linear(self.hidden_dim, self.out_dim))
        self.anticipation_layer = nn.Sequential(nn.Linear(self.hidden_dim, self.anticipation_length * self.hidden_dim))

        self.precision = cfg.get("precision", "fp16")
        self.train_precision = cfg.get("train_precision", "tf32x3")
        self.chunk_frames = int(cfg.get("chunk_frames", 1 << 17))
        self._handle = None
        self._handle_device = None
        self._packed_key = None
        self._packed_formats = 0
        self._ant_key = None
        self._workspace = None
        self._ant_workspace = None
        self.last_labels = None
        self.last_anticipation_labels = None  # int32 [B, T, A]
        self.anticipation_slab_rows = int(cfg.get("anticipation_slab_rows", 16384))

    def _release(self):
        super()._release()
        self._ant_key = None

    def _sync_weights(self, lib, device, formats=_lib.PACK_ALL):
        super()._sync_weights(lib, device, formats)
        lin = self.anticipation_layer[0]
        key = (lin.weight.data_ptr(), lin.weight._version, lin.bias.data_ptr(), lin.bias._version)
        if key == self._ant_key:
            return
        for k, t in zip(_ANT_KEYS, (lin.weight, lin.bias)):
            if t.device != device or t.dtype != torch.float32:
                raise RuntimeError(f"parameter {k} must be fp32 on {device} (got {t.dtype} on {t.device}); call model.to(device)")
        w, b = lin.weight.detach().contiguous(), lin.bias.detach().contiguous()
        stream = torch.cuda.current_stream(device).cuda_stream
        _lib.check(lib.prego_model_load_anticipation(self._handle, self.anticipation_length, w.data_ptr(), b.data_ptr(), stream),
                   "prego_model_load_anticipation")
        self._ant_key = key

    def online_session(self, *args, **kwargs):
        raise RuntimeError("per-frame online sessions serve the MROAD head only; use MROADA.infer(..., want_anticipation=True)")

    def forward(self, rgb_input, flow_input):
        """rnn.py:112-137, eval mode."""
        if self.training:
            raise RuntimeError("prego_b200.MROADA is inference-only (call model.eval()); training covers MiniROAD (MROAD)")
        out = self.infer(rgb_input, flow_input, want_probs=True, want_labels=True, want_anticipation=True)
        self.last_labels = out["labels"]
        self.last_anticipation_labels = out["anticipation_labels"]
        return {"logits": out["probs"], "anticipation_logits": out["anticipation_probs"]}


class OnlineSession:
    """One frame per call for ``B`` <= 8 concurrent streams (GEMV kernels of online_kernels.cuh, replayed as one
    CUDA graph per frame through ``prego_online_*``).  Buffers are allocated once; ``step`` patches two pointers.
    The packed weights are captured when the session is opened: open a new session after changing parameters."""

    def __init__(self, model: MROAD, num_streams: int, device=None, precision=None, want_probs=False, host_labels=False):
        lib = _lib.load()
        if device is None:
            device = next(model.parameters()).device
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("prego_b200 runs on CUDA (sm_100a) only; there is no CPU fallback")
        prec_name = precision or model.precision
        if prec_name == "fp32":
            raise RuntimeError("online sessions run in 'fp16' or 'bf16' (use MROAD.infer for the exact fp32 path)")
        self.model, self.device, self.B = model, device, int(num_streams)
        self._lib, self._session = lib, C.c_void_p()
        with torch.cuda.device(device):
            model._ensure_handle(device)
            model._sync_weights(lib, device)
            self.h = torch.zeros(self.B, model.hidden_dim, dtype=torch.float32, device=device)
            # host_labels: the kernel stores the labels straight into pinned (device-mapped) host memory, so the caller
            # reads them after a stream sync without a D2H copy on the per-frame critical path
            self.labels = (torch.zeros(self.B, 1, dtype=torch.int32).pin_memory() if host_labels
                           else torch.empty(self.B, 1, dtype=torch.int32, device=device))
            self.labels_np = self.labels.numpy() if host_labels else None
            self.probs = torch.empty(self.B, 1, model.out_dim, dtype=torch.float32, device=device) if want_probs else None
            self.stream = torch.cuda.current_stream(device)
            self._stream_ptr = self.stream.cuda_stream
            self._step_wait = lib.prego_online_step_wait
            self.stream.synchronize()  # weight packing done before the graph is built
            _lib.check(lib.prego_online_open(model._handle, self.B, _lib.PRECISIONS[prec_name], self.h.data_ptr(),
                                             self.probs.data_ptr() if want_probs else None, None, self.labels.data_ptr(),
                                             C.byref(self._session)), "prego_online_open")

    def step_wait(self, rgb_frame, flow_frame):
        """``step`` + host-side completion without a stream synchronize (``prego_online_step_wait``): returns once the
        frame's labels are readable.  For ``host_labels`` sessions this is the whole per-frame round trip; the labels
        are also exposed as the numpy view ``labels_np`` (no torch dispatch on the per-frame path).  One foreign call per
        frame, launched on the stream that was current when the session was opened (``self.stream``; looking the current
        stream up through torch costs more than a microsecond per frame)."""
        rc = self._step_wait(self._session, rgb_frame.data_ptr() if rgb_frame is not None else None,
                             flow_frame.data_ptr() if flow_frame is not None else None, self._stream_ptr)
        if rc:
            _lib.check(rc, "prego_online_step_wait")
        return self.labels

    def close(self):
        if self._session is not None and self._session.value:
            self._lib.prego_online_close(self._session)
            self._session = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self):
        self.h.zero_()

    def trace(self):
        """Per-CTA SM-clock stamps [ctas, 16] of the last frame's phase boundaries (PREGO_ONLINE_TRACE=1 sessions)."""
        import numpy as np
        out = np.zeros((256, 16), dtype=np.int64)
        n = C.c_int32(0)
        _lib.check(self._lib.prego_online_trace(self._session, out.ctypes.data, 256, C.byref(n)), "prego_online_trace")
        return out[:n.value]

    def step(self, rgb_frame, flow_frame):
        """rgb_frame / flow_frame: contiguous fp32 CUDA tensors with B * D elements ([B, D] or [B, 1, D]).
        Returns the int32 label tensor [B, 1] (device, or pinned host memory for ``host_labels`` sessions: valid
        after the stream is synchronized; overwritten by the next step)."""
        rc = self._lib.prego_online_step(self._session, rgb_frame.data_ptr() if rgb_frame is not None else None,
                                         flow_frame.data_ptr() if flow_frame is not None else None,
                                         torch.cuda.current_stream(self.device).cuda_stream)
        if rc:
            _lib.check(rc, "prego_online_step")
        return self.labels

"""Feature ingest for the hot path (reference: ``step_recognition/datasets/dataset.py:45-95,120-132``).

The reference loader keeps every video's TSN features as float64/float32 ``.npy`` arrays, converts them to fp32 per
item, and -- in both shipped configs -- replaces the flow stream by ``np.zeros`` (``dataset.py:63-69``).  End to end
the B200 path is bound by the host->device link (16 KiB per frame as fp32), so the ingest side does four things:

* **convert once**: each video is rounded to the 16-bit operand format of the model's precision when it is loaded
  (the CUDA path would apply exactly this rounding in its staging pass, so results are bit-identical) and kept in
  pinned host memory: 4 KiB per frame and stream instead of 8;
* **drop the zero flow**: a video whose flow stream is the reference's all-zero dummy keeps no flow array at all; the
  model is told so (``zero_flow``) and skips that half of the projection;
* **stream**: videos are bucketed by length (the GRU is causal, so end-padding cannot change earlier outputs), each
  bucket is assembled in a pinned staging buffer and copied on a side stream while the previous bucket computes.

* **round on the host when the caller holds fp32 host tensors** (what the reference's ``DataLoader`` yields,
  ``dataset_builder.py:17-23``): ``HostRoundingStager`` applies the device's operand rounding on host threads and copies
  through a cache-resident pinned ring (``csrc/host_stage.cpp``), halving the bytes on the link with bit-identical
  results; the loader's all-zero flow dummy is recognised by a host scan and never copied.

Host code only; the compute stays behind ``MROAD.infer`` (there is no CPU fallback here either).
"""
from __future__ import annotations

import json
import os
import os.path as osp
from dataclasses import dataclass
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import ctypes as C
import threading

import numpy as np
import torch

from . import _lib

OPERAND_DTYPES = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}
_F16_MAX = 65504.0


def to_operand(a, dtype: torch.dtype, pin: bool = False) -> torch.Tensor:
    """numpy / torch array of any float type -> contiguous host tensor in ``dtype`` with the device path's rounding:
    round-to-nearest-even from fp32, fp16 saturating at +-65504 (Op16<0>::pack2 in csrc/gemm_tc.cuh)."""
    t = torch.as_tensor(np.ascontiguousarray(a)) if not isinstance(a, torch.Tensor) else a
    t = t.to(torch.float32)  # the reference's own first step (dataset.py:129-131)
    if dtype == torch.float16:
        t = t.clamp(-_F16_MAX, _F16_MAX)
    t = t.to(dtype).contiguous()
    if pin and torch.cuda.is_available():
        t = t.pin_memory()
    return t


@dataclass
class Video:
    vid: str
    rgb: Optional[torch.Tensor]   # [T, Dr] host, operand dtype (None when the model runs flow-only)
    flow: Optional[torch.Tensor]  # [T, Df] host, or None = all-zero flow (dataset.py:63-69)
    gt: Optional[np.ndarray]      # int labels [T] (argmax of the one-hot targets), if annotations were given

    @property
    def T(self) -> int:
        return int((self.rgb if self.rgb is not None else self.flow).shape[0])


class FeatureStore:
    """All videos of an evaluation set, converted once and pinned."""

    def __init__(self, videos: Sequence[Video], dtype: torch.dtype):
        self.videos, self.dtype = list(videos), dtype

    def __len__(self):
        return len(self.videos)

    @property
    def zero_flow(self) -> bool:
        return all(v.flow is None for v in self.videos)

    @property
    def frames(self) -> int:
        return sum(v.T for v in self.videos)

    def host_bytes_per_frame(self) -> float:
        b = sum((0 if v.rgb is None else v.rgb.numel() * v.rgb.element_size()) + (0 if v.flow is None else v.flow.numel() * v.flow.element_size())
                for v in self.videos)
        return b / max(self.frames, 1)

    @classmethod
    def from_arrays(cls, items: Sequence[Tuple[str, object, object, object]], precision: str = "fp16", pin: bool = True) -> "FeatureStore":
        """items: (vid, rgb[T, Dr], flow[T, Df] | None, target[T, K] one-hot | labels[T] | None).  A flow array that is
        all zero is dropped like the reference's dummy."""
        dtype = OPERAND_DTYPES[precision]
        vids = []
        for vid, rgb, flow, target in items:
            if flow is not None:
                fz = not bool(np.any(flow)) if isinstance(flow, np.ndarray) else not bool(torch.any(flow != 0))
                flow = None if fz else flow
            gt = None
            if target is not None:
                tg = np.asarray(target)
                gt = tg.argmax(axis=1) if tg.ndim == 2 else tg.astype(np.int64)
            vids.append(Video(vid, None if rgb is None else to_operand(rgb, dtype, pin), None if flow is None else to_operand(flow, dtype, pin), gt))
        return cls(vids, dtype)

    @classmethod
    def from_reference_layout(cls, cfg: dict, precision: str = "fp16", split: str = "test_session_set", pin: bool = True) -> "FeatureStore":
        """The reference's on-disk layout (dataset.py:45-95): ``<root>/<rgb_type>/<vid>.npy``, annotations under
        ``<root>/<annotation_type>/<vid>.npy``, flow under the BNInception path unless ``flow_type`` is the
        ``flow_anet_resnet50`` dummy (all zero -> not stored).  Unreadable videos are skipped like dataset.py:87-93."""
        root = cfg["root_path"]
        names = json.load(open(cfg["video_list_path"]))[cfg["data_name"]][split]
        items = []
        for vid in names:
            try:
                target = np.load(osp.join(root, cfg["annotation_type"], vid + ".npy"))
                rgb = None if cfg.get("no_rgb") else np.load(osp.join(root, cfg["rgb_type"], vid + ".npy"))
                if cfg.get("no_flow") or cfg["flow_type"] == "flow_anet_resnet50":
                    flow = None
                else:
                    flow = np.load(osp.join(root, cfg["flow_type"], "assembly_optical_flow_BNInception", vid, "assembling.npy"))
                items.append((vid, rgb, flow, target))
            except Exception as e:
                print("---- Exception in loading video ", e)
        return cls.from_arrays(items, precision, pin)


def bucket_order(lengths: Sequence[int], batch_streams: int) -> List[List[int]]:
    """Indices grouped into batches of at most ``batch_streams`` videos, longest first (padding waste is the difference
    to the longest video of a batch, so neighbours in length share a batch)."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    return [order[s:s + batch_streams] for s in range(0, len(order), batch_streams)]


def stream_batches(store: FeatureStore, device, batch_streams: int = 64) -> Iterator[Tuple[List[int], torch.Tensor, Optional[torch.Tensor]]]:
    """Yields ``(indices, rgb[B, Tmax, Dr], flow[B, Tmax, Df] | None)`` device tensors (operand dtype, end-padded with
    zeros), pipelined over three slots: while the consumer works on batch i, batch i + 1 is already on the device and
    batch i + 2 is assembled in pinned memory and copied on a side stream.  The yielded tensors stay valid until the
    next batch is requested."""
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("prego_b200 ingest streams to a CUDA device; there is no CPU path")
    batches = bucket_order([v.T for v in store.videos], batch_streams)
    if not batches:
        return
    use_flow = not store.zero_flow
    v0 = store.videos[0]
    dr = 0 if v0.rgb is None else int(v0.rgb.shape[1])
    df = 0
    for v in store.videos:
        if v.flow is not None:
            df = int(v.flow.shape[1])
            break
    tmax_all = store.videos[batches[0][0]].T
    bmax = max(len(b) for b in batches)

    def alloc(d):
        if d == 0:
            return None, None
        # flat buffers: every batch views them as a DENSE [B, Tmax_of_batch, d] tensor (what the C ABI wants)
        n = bmax * tmax_all * d
        return torch.zeros(n, dtype=store.dtype).pin_memory(), torch.empty(n, dtype=store.dtype, device=device)

    # three slots: while the consumer computes on batch i, batch i + 1 is already on the device and batch i + 2 is being
    # assembled into the slot batch i - 1 used (whose compute was launched one yield ago)
    NS = 3
    slots = []
    for _ in range(NS):
        hr, drt = alloc(dr)
        hf, dft = alloc(df if use_flow else 0)
        slots.append({"hr": hr, "dr": drt, "hf": hf, "df": dft, "ready": torch.cuda.Event(), "free": torch.cuda.Event(), "tmax": 0})
    copy_stream = torch.cuda.Stream(device=device)
    main = torch.cuda.current_stream(device)
    for s in slots:
        s["free"].record(main)

    def fill(slot, idx):
        tmax = store.videos[idx[0]].T
        slot["free"].synchronize()  # the consumer's work on this slot's previous batch has finished
        for name, attr, d in (("hr", "rgb", dr), ("hf", "flow", df)):
            if slot[name] is None:
                continue
            h = slot[name][: len(idx) * tmax * d].view(len(idx), tmax, d)
            for j, i in enumerate(idx):
                src = getattr(store.videos[i], attr)
                t = 0 if src is None else int(src.shape[0])
                if t:
                    h[j, :t].copy_(src)
                h[j, t:].zero_()
        with torch.cuda.stream(copy_stream):
            for hn, dn, d in (("hr", "dr", dr), ("hf", "df", df)):
                if slot[hn] is not None:
                    n = len(idx) * tmax * d
                    slot[dn][:n].copy_(slot[hn][:n], non_blocking=True)
            slot["ready"].record(copy_stream)
        slot["tmax"] = tmax

    for bi in range(min(2, len(batches))):
        fill(slots[bi % NS], batches[bi])
    for bi, idx in enumerate(batches):
        slot = slots[bi % NS]
        tmax = slot["tmax"]
        main.wait_event(slot["ready"])
        rgb = None if slot["dr"] is None else slot["dr"][: len(idx) * tmax * dr].view(len(idx), tmax, dr)
        flow = None if slot["df"] is None else slot["df"][: len(idx) * tmax * df].view(len(idx), tmax, df)
        yield idx, rgb, flow
        slot["free"].record(main)
        if bi + 2 < len(batches):
            fill(slots[(bi + 2) % NS], batches[bi + 2])


@torch.no_grad()
def predict_labels_streamed(model, store: FeatureStore, device, batch_streams: int = 64, precision: Optional[str] = None) -> Dict[str, torch.Tensor]:
    """{vid: int32 labels[T]} (device) for every video of the store, fed through ``stream_batches``."""
    model.eval()
    prec = precision or model.precision
    if OPERAND_DTYPES[prec] != store.dtype:
        raise RuntimeError(f"store holds {store.dtype} features but the model runs in '{prec}'")
    out: Dict[str, torch.Tensor] = {}
    zero_flow = store.zero_flow and model.use_flow
    for idx, rgb, flow in stream_batches(store, device, batch_streams):
        labels = model.infer(rgb, flow, want_probs=False, want_labels=True, precision=prec, zero_flow=zero_flow)["labels"]
        for j, i in enumerate(idx):
            v = store.videos[i]
            out[v.vid] = labels[j, : v.T].clone()
    if hasattr(model, "check_device"):
        model.check_device("predict_labels_streamed")
    return {v.vid: out[v.vid] for v in store.videos}


def round_features_host(src: torch.Tensor, dst: torch.Tensor, precision: str = "fp16", threads: Optional[int] = None) -> torch.Tensor:
    """fp32 host tensor -> the 16-bit operand format of ``precision`` in ``dst`` (host, same numel) with the device
    path's rounding, on ``threads`` host threads (``prego_host_round_features``; releases the GIL)."""
    if src.is_cuda or dst.is_cuda or src.dtype != torch.float32 or dst.dtype != OPERAND_DTYPES[precision] or precision == "fp32":
        raise RuntimeError("round_features_host: fp32 host source and fp16 / bf16 host destination expected")
    if not src.is_contiguous() or not dst.is_contiguous() or src.numel() != dst.numel():
        raise RuntimeError("round_features_host: contiguous tensors of equal size expected")
    _lib.check(_lib.load().prego_host_round_features(src.data_ptr(), dst.data_ptr(), src.numel(), _lib.PRECISIONS[precision],
                                                     int(threads or os.cpu_count() or 1)), "prego_host_round_features")
    return dst


class HostRoundingStager:
    """fp32 HOST feature batches -> device tensors in the 16-bit operand format, with the rounding done on the host.

    The reference hands the model fp32 host tensors (``datasets/dataset.py:120-132``, ``trainer/eval.py:40-42``); end to end
    the path is bound by the host->device link, and the first thing the device does with a feature is round it to the
    projection GEMM's operand format.  Rounding on the host with the same rule (``prego_host_round_features``) halves
    the bytes on the link; ``MROAD.infer`` then reads the tensors in place (``PREGO_FEAT_16``) and returns bit-identical
    results.  The rounded values travel through a small pinned ring (``ring_slots`` x ``ring_slot_bytes``): a slot is
    copied (copy stream) while the next one is rounded (persistent pool of host threads claiming 128 KiB chunks), and
    it is still in the CPU's last-level cache when the DMA engine reads it, so the host's DRAM only sees the fp32 read.  The whole batch i + 1 is
    staged while the device computes batch i (two device slots).

    Rounding costs host memory bandwidth (4 B read + 2 B written per value, then 2 B read by the DMA engine), and on a
    host whose memory system is the limit the link idles part of the time.  ``direct_streams`` > 0 sends that many
    leading streams as plain fp32 (no host work, 16 KiB per frame on the link) so that link and host finish together;
    the device then runs the batch as two calls (fp32 rows, 16-bit rows) -- still bit-identical, every stream is
    independent.

        stager = HostRoundingStager(B, T, d_rgb, d_flow, "fp16", device)
        stager.submit(0, rgb_host, flow_host)                # returns at once; a worker thread rounds and copies
        labels = stager.infer(model, 0, h_state=h)["labels"] # waits for the copies on the current stream, runs, releases
    """

    def __init__(self, B: int, T: int, d_rgb: int, d_flow: int, precision: str, device, threads: Optional[int] = None,
                 slots: int = 2, direct_streams: int = 0, ring_slots: int = 3, ring_slot_bytes: int = 8 << 20,
                 detect_zero_flow: bool = True):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("prego_b200 ingest streams to a CUDA device; there is no CPU path")
        if precision not in ("fp16", "bf16"):
            raise RuntimeError("host rounding targets the 16-bit operand formats ('fp16' / 'bf16')")
        self.precision, self.dtype = precision, OPERAND_DTYPES[precision]
        self.B, self.T, self.dims = int(B), int(T), (int(d_rgb), int(d_flow))
        self.Bd = max(0, min(int(direct_streams), self.B))
        self.detect_zero_flow = bool(detect_zero_flow) and self.dims[0] > 0
        cores = os.cpu_count() or 1
        self.threads = int(threads or (cores - 2 if cores > 4 else cores))  # the rounding is memory-bound: leave two cores to Python / the driver
        self.copy_stream = torch.cuda.Stream(device=self.device)
        # the plain-fp32 rows of a split batch travel on their OWN stream: behind the ring's small copies on one stream they would
        # serialise with them and stall the rounding ring until the large copy is through
        self.direct_stream = torch.cuda.Stream(device=self.device) if self.Bd else None
        self.slots = []
        B16 = self.B - self.Bd
        # the C ring stager: a persistent pool of rounding threads + a pinned ring small enough to stay in the CPU's
        # last-level cache between the rounding stores and the DMA read (csrc/host_stage.cpp)
        self._ring = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(_lib.load().prego_host_stager_create(self.threads, int(ring_slots), int(ring_slot_bytes), C.byref(self._ring)),
                       "prego_host_stager_create")
        for _ in range(slots):
            dev = [torch.empty(B16, self.T, d, dtype=self.dtype, device=self.device) if d and B16 else None for d in self.dims]
            dev32 = [torch.empty(self.Bd, self.T, d, dtype=torch.float32, device=self.device) if d and self.Bd else None for d in self.dims]
            self.slots.append({"dev": dev, "dev32": dev32, "ready": torch.cuda.Event(), "free": torch.cuda.Event(), "direct_done": torch.cuda.Event(),
                               "issued": threading.Event(), "thread": None, "error": None})
        main = torch.cuda.current_stream(self.device)
        for s in self.slots:
            s["free"].record(main)
            s["ready"].record(main)

    def _work(self, slot, srcs):
        try:
            lib = _lib.load()
            prec = _lib.PRECISIONS[self.precision]
            if self.detect_zero_flow and srcs[1] is not None and lib.prego_host_all_zero(srcs[1].data_ptr(), srcs[1].numel(), self.threads) == 1:
                # the reference loader's flow dummy (np.zeros, datasets/dataset.py:63-69): neither rounded nor copied; infer()
                # declares it (flow_is_zero: bit-identical to multiplying by the zeros)
                srcs = [srcs[0], None]
                slot["flow"], slot["zero_flow"] = False, True
            with torch.cuda.device(self.device), torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(slot["free"])  # the consumer is done with this slot's device tensors
                if self.direct_stream is not None:
                    self.direct_stream.wait_event(slot["free"])
                    with torch.cuda.stream(self.direct_stream):  # fp32 rows: a second DMA queue next to the ring's copies
                        for src, d32 in zip(srcs, slot["dev32"]):
                            if src is not None and d32 is not None:
                                d32.copy_(src[:self.Bd], non_blocking=True)
                        slot["direct_done"].record(self.direct_stream)
                for src, dev in zip(srcs, slot["dev"]):
                    if src is None or dev is None:
                        continue
                    _lib.check(lib.prego_host_stager_run(self._ring, src[self.Bd:].data_ptr(), dev.data_ptr(), dev.numel(), prec,
                                                         self.copy_stream.cuda_stream), "prego_host_stager_run")
                if self.direct_stream is not None:
                    self.copy_stream.wait_event(slot["direct_done"])
                slot["ready"].record(self.copy_stream)
        except Exception as e:  # surfaced by wait()
            slot["error"] = e
        finally:
            slot["issued"].set()

    def submit(self, i: int, rgb_host: Optional[torch.Tensor], flow_host: Optional[torch.Tensor]) -> None:
        slot = self.slots[i % len(self.slots)]
        srcs = []
        for x, d in zip((rgb_host, flow_host), self.dims):
            if x is None or d == 0:
                srcs.append(None)
                continue
            if x.is_cuda or x.dtype != torch.float32 or tuple(x.shape) != (self.B, self.T, d) or not x.is_contiguous():
                raise RuntimeError(f"expected a contiguous fp32 host tensor [{self.B}, {self.T}, {d}], got {x.dtype} {tuple(x.shape)}")
            srcs.append(x)
        if slot["thread"] is not None:
            slot["thread"].join()
        slot["issued"].clear()
        slot["error"] = None
        slot["flow"], slot["zero_flow"] = srcs[1] is not None, False
        slot["thread"] = threading.Thread(target=self._work, args=(slot, srcs), daemon=True)
        slot["thread"].start()

    def wait(self, i: int):
        """-> ((rgb32, flow32 | None) of the first ``direct_streams`` streams or None, (rgb16, flow16 | None) of the rest
        or None); the current stream waits for their copies."""
        slot = self.slots[i % len(self.slots)]
        if slot["thread"] is None:
            raise RuntimeError(f"wait({i}) without a matching submit({i})")
        slot["issued"].wait()
        if slot["error"] is not None:
            raise slot["error"]
        torch.cuda.current_stream(self.device).wait_event(slot["ready"])
        flow = bool(slot.get("flow"))
        g32 = (slot["dev32"][0], slot["dev32"][1] if flow else None) if self.Bd else None
        g16 = (slot["dev"][0], slot["dev"][1] if flow else None) if self.B > self.Bd else None
        return g32, g16

    def release(self, i: int) -> None:
        """The consumer's work on submission ``i`` has been enqueued: its device tensors may be overwritten after it."""
        self.slots[i % len(self.slots)]["free"].record(torch.cuda.current_stream(self.device))

    def infer(self, model, i: int, h_state: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None, **kw):
        """wait(i) + ``model.infer`` on the staged tensors + release(i).  Returns {'labels': int32 [B, T]} (``labels`` is
        reused when given).  ``h_state`` [B, H] is carried in place.  kw: chunk_T, zero_flow (set automatically when the
        submitted flow tensor was recognised as the all-zero dummy)."""
        g32, g16 = self.wait(i)
        if self.slots[i % len(self.slots)].get("zero_flow"):
            kw = dict(kw, zero_flow=True)
        if labels is None:
            labels = torch.empty(self.B, self.T, dtype=torch.int32, device=self.device)
        for grp, b0, b1 in ((g32, 0, self.Bd), (g16, self.Bd, self.B)):
            if grp is None:
                continue
            model.infer(grp[0], grp[1], h_state=None if h_state is None else h_state[b0:b1], want_probs=False,
                        precision=self.precision, out={"labels": labels[b0:b1]}, **kw)
        self.release(i)
        return {"labels": labels}

    def close(self):
        for s in self.slots:
            if s["thread"] is not None:
                s["thread"].join()
                s["thread"] = None
        if self._ring is not None and self._ring.value:
            torch.cuda.synchronize(self.device)
            _lib.load().prego_host_stager_destroy(self._ring)
            self._ring = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""Device-resident restatement of the per-batch evaluation body of the reference engine,
``engine/crog_engine.py:386-556`` (``inference_with_grasp``).  ``GraspEvaluator.step`` is the
synthetic-benchmark contract of SURVEY.md §8(d) (maps decoded at the network resolution, identity
``ori_size``); ``GraspEvaluator.step_original`` is the real pipeline, with the OpenCV-exact inverse
letterbox warp to the original image size (row f-1) between the glue and the decode, plus the
mask-IoU / Pr@K bookkeeping.

    model(img, word) -> sigmoid(mask, qua, wid) + bicubic x4 (align_corners=True)   [crog_engine.py:446-474]
                     -> detect_grasps(K=1 and K=5) per sample                      [:519-522]
                     -> calculate_jacquard_index against the sample's GT           [:524-527]
                     -> correct/total counters for J@1 and J@5                     [:526-527,535-537]

Everything between the model and the counters stays in HBM; the reference instead copies ten
maps per sample to the host and loops in Python.  J@1 uses the first of the top-5 peaks: the
greedy peak order does not depend on K, so ``detect_grasps(K=1)`` is the first row of K=5.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

from . import _lib as L
from .utils import grasp_eval as GE
from .utils import warp as WP

SIGMOID_PLANES = 0b10011  # mask, qua, wid get a sigmoid; sin, cos stay raw (crog_engine.py:446-448)


def postprocess(maps: Sequence[torch.Tensor], size: Tuple[int, int]) -> torch.Tensor:
    """5 logits maps B x 1 x h x w (any contiguous layout) -> tensor [5, B, H, W] of sigmoid + bicubic maps."""
    lib = L.lib()
    stacked = maps if torch.is_tensor(maps) else torch.stack([m.reshape(m.shape[0], m.shape[-2], m.shape[-1]) for m in maps])
    stacked = stacked.reshape(len(maps), -1, stacked.shape[-2], stacked.shape[-1]).contiguous().float()
    NP, B, h, w = stacked.shape
    out = torch.empty((NP, B, size[0], size[1]), dtype=torch.float32, device=stacked.device)
    with torch.cuda.device(stacked.device):
        L.check(lib.crog_sigmoid_bicubic(stacked.data_ptr(), out.data_ptr(), NP, B, h, w, size[0], size[1],
                                         SIGMOID_PLANES if NP == 5 else 0b1, L.stream_ptr()))
    return out


class GraspEvaluator:
    """Accumulates J@1 / J@5 over batches on the device; ``reduce()`` sums the int64[4] counters over ranks
    (one all-reduce per evaluation — the reference leaves the per-rank counters un-reduced)."""

    def __init__(self, model, num_grasps: int = 5, device: Optional[torch.device] = None):
        self.model = model
        self.K = num_grasps
        dev = device or next(model.parameters()).device
        self.counters = torch.zeros(4, dtype=torch.int64, device=dev)
        self.iou_list = []  # per-batch device tensors of mask IoU (crog_engine.py:518-519)

    @torch.no_grad()
    def step_original(self, img: torch.Tensor, word: torch.Tensor, gt: torch.Tensor, gt_count: torch.Tensor, inverse_mats,
                      ori_size: Tuple[int, int], mask_target: Optional[torch.Tensor] = None, mask_thr: float = 0.35):
        """crog_engine.py:429-527 for a batch whose images share one original size ``ori_size = (h, w)``:
        ``inverse_mats`` are the dataset's ``inverse`` matrices (2x3, one or B of them), ``mask_target`` the
        letterboxed GT masks [B,S,S] (optional).  All five maps are warped back with cv2.warpAffine semantics
        (INTER_CUBIC, borderValue 0), the mask is thresholded at 0.35 for IoU, the other four are decoded.
        Returns a dict: post [5,B,S,S], maps [5,B,h,w], iou [B] | None, peaks, n_peaks, grasps, j_flags."""
        post = postprocess(self._logits(img, word), (img.shape[-2], img.shape[-1]))
        h, w = int(ori_size[0]), int(ori_size[1])
        inv = WP.warp_affine_cubic(post, inverse_mats, (w, h), 0.0)
        iou = None
        if mask_target is not None:
            tgt = WP.warp_affine_cubic(mask_target.reshape(mask_target.shape[0], mask_target.shape[-2], mask_target.shape[-1]),
                                       inverse_mats, (w, h), 0.0)
            iou, _ = WP.mask_iou(inv[0], tgt, mask_thr)
            self.iou_list.append(iou)
        peaks, n, grasps = GE.detect_grasps_batched(inv[1], inv[2], inv[3], inv[4], self.K)
        flags = GE.jacquard_batched(grasps, n, gt, gt_count, counters=self.counters)
        return {"post": post, "maps": inv, "iou": iou, "peaks": peaks, "n_peaks": n, "grasps": grasps, "j_flags": flags}

    def gathered_iou(self) -> Optional[torch.Tensor]:
        """Per-sample mask IoU of every rank, concatenated in rank order — the reference's ``concat_all_gather(iou_list)``
        (utils/misc.py:47-59, engine/crog_engine.py:269).  Unlike the reference's fixed-shape all_gather, ragged shards
        (shard_range gives the low ranks one sample more) are handled by gathering the counts first."""
        import torch.distributed as dist

        if not self.iou_list:
            return None
        iou = torch.cat(self.iou_list)
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return iou
        world = dist.get_world_size()
        n = torch.tensor([iou.numel()], dtype=torch.int64, device=iou.device)
        counts = [torch.zeros_like(n) for _ in range(world)]
        dist.all_gather(counts, n)
        counts = [int(c.item()) for c in counts]
        pad = torch.zeros(max(counts), dtype=iou.dtype, device=iou.device)
        pad[:iou.numel()] = iou
        parts = [torch.zeros_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad)
        return torch.cat([p[:c] for p, c in zip(parts, counts)])

    def summary(self, gather: bool = True):
        """crog_engine.py:535-556: mean mask IoU, Pr@50..90 and J@1 / J@K over everything seen so far.  With ``gather`` and
        an initialised process group the IoU list is all-gathered over ranks first, as the reference does
        (crog_engine.py:269); the J counters are whatever ``reduce()`` has made of them (per rank until it is called)."""
        j1, jk = self.j_index()
        out = {"J@1": j1, f"J@{self.K}": jk, "n": 0, "IoU": float("nan"), "Pr": {}}
        iou = self.gathered_iou() if gather else (torch.cat(self.iou_list) if self.iou_list else None)
        if iou is not None:
            out["n"], out["IoU"] = int(iou.numel()), float(iou.mean().item())
            for t in range(5, 10):
                thr = torch.arange(0.5, 1.0, 0.1)[t - 5].item()  # the reference's float32 thresholds (crog_engine.py:541)
                out["Pr"][f"Pr@{t * 10}"] = float((iou > thr).float().mean().item())
        return out

    @torch.no_grad()
    def step(self, img: torch.Tensor, word: torch.Tensor, gt: torch.Tensor, gt_count: torch.Tensor):
        """img B x 3 x S x S, word B x L, gt B x M x 6 float64 (edited in place like the reference), gt_count B int32.
        Returns (post [5,B,S,S], peaks, n_peaks, grasps, j_flags)."""
        post = postprocess(self._logits(img, word), (img.shape[-2], img.shape[-1]))
        peaks, n, grasps = GE.detect_grasps_batched(post[1], post[2], post[3], post[4], self.K)
        flags = GE.jacquard_batched(grasps, n, gt, gt_count, counters=self.counters)
        return post, peaks, n, grasps, flags

    def _logits(self, img, word):
        """The five logits maps as one [5, B, 1, h, w] tensor when the model offers it (no re-stacking), else its tuple."""
        fs = getattr(self.model, "forward_stacked", None)  # wrapped models (DataParallel) only have forward
        return fs(img, word) if fs is not None else self.model(img, word)[0]

    @torch.no_grad()
    def stream(self, host_batches, letterbox=None, original: bool = False):
        """End-to-end evaluation over HOST batches with the input copies hidden behind the previous batch's compute.

        ``host_batches`` yields ``(img, word, gt, gt_count)`` CPU tensors (pinned memory for truly asynchronous copies).
        ``img`` is either the float32 network input [B,3,S,S] or — with ``letterbox = (mat, mat_inv, (ori_h, ori_w))``
        from ``utils.warp.get_transform_mat`` — the raw uint8 RGB camera frames [B,ori_h,ori_w,3]: they cross PCIe as
        bytes (0.92 MB instead of 2.08 MB per sample) and the letterbox + normalisation of utils/dataset.py:843-866 runs
        on the device (``crog_preprocess_u8``, bit-exact with OpenCV).  ``original=True`` additionally decodes at the
        original image size after the inverse letterbox (``step_original``, the reference's real evaluation loop,
        engine/crog_engine.py:386-556) instead of at the network resolution.

        Two device staging slots are filled on a copy stream: while batch k runs on the compute stream, batch k+1 is
        already crossing PCIe.  For every batch the decoded grasps / peak counts / J flags are copied back and the
        host waits for them (the caller reads the result of every step); yields ``(n_peaks, grasps, j_flags)`` as
        pinned CPU tensors that stay valid until the next-but-one iteration.  The launches of batch k+1 are enqueued
        before the host blocks on the results of batch k, so the GPU does not idle while the host hands a result over."""
        dev = self.counters.device
        main = torch.cuda.current_stream(dev)
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._slots = [None, None]
            self._ready = [torch.cuda.Event(), torch.cuda.Event()]
            self._free = [torch.cuda.Event(), torch.cuda.Event()]
            self._done = [torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()]
            self._hout = [None, None, None]
            self._affine = {}
        if original and letterbox is None:
            raise RuntimeError("stream(original=True) needs letterbox=(mat, mat_inv, ori_size)")
        it = iter(host_batches)

        def prefetch(k):
            try:
                hb = next(it)
            except StopIteration:
                return False
            s = k & 1
            if self._slots[s] is None or any(d.shape != h.shape or d.dtype != h.dtype for d, h in zip(self._slots[s], hb)):
                self._slots[s] = [torch.empty(h.shape, dtype=h.dtype, device=dev) for h in hb]
            with torch.cuda.stream(self._copy_stream):
                if k >= 2:
                    self._copy_stream.wait_event(self._free[s])  # batch k-2 no longer reads this slot
                for d, h in zip(self._slots[s], hb):
                    d.copy_(h, non_blocking=True)
                self._ready[s].record(self._copy_stream)
            return True

        def affine(mat, B):
            key = (id(mat), B)
            if key not in self._affine:
                if len(self._affine) > 8:
                    self._affine.clear()
                self._affine[key] = WP.device_affine(mat, B, dev)
            return self._affine[key]

        k, more = 0, prefetch(0)
        pending = None  # result slot of the batch whose launches are enqueued but whose results are not handed over yet
        while more:
            s = k & 1
            more = prefetch(k + 1)
            main.wait_event(self._ready[s])
            img, word, gt, cnt = self._slots[s]
            if img.dtype == torch.uint8:
                if letterbox is None:
                    raise RuntimeError("uint8 frames need letterbox=(mat, mat_inv, ori_size)")
                S = self.model.cfg.input_size
                buf = self.model.input_buffer(img.shape[0], S) if hasattr(self.model, "input_buffer") else None
                img = WP.preprocess_images(img, affine(letterbox[0], img.shape[0]), (S, S), out=buf)
            if original:
                r_ = self.step_original(img, word, gt, cnt, affine(letterbox[1], img.shape[0]), letterbox[2])
                n, grasps, flags = r_["n_peaks"], r_["grasps"], r_["j_flags"]
            else:
                _, _, n, grasps, flags = self.step(img, word, gt, cnt)
            self._free[s].record(main)
            r = k % 3
            if self._hout[r] is None or self._hout[r][1].shape != grasps.shape:
                self._hout[r] = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in (n, grasps, flags)]
            for h, d in zip(self._hout[r], (n, grasps, flags)):
                h.copy_(d, non_blocking=True)
            self._done[r].record(main)
            if pending is not None:
                self._done[pending].synchronize()
                yield tuple(self._hout[pending])
            pending = r
            k += 1
        if pending is not None:
            self._done[pending].synchronize()
            yield tuple(self._hout[pending])

    def reduce(self) -> torch.Tensor:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.counters, op=dist.ReduceOp.SUM)
        return self.counters

    def j_index(self):
        c = self.counters.tolist()
        return (c[0] / max(c[1], 1), c[2] / max(c[3], 1))


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous shard of ``n`` samples for ``rank``; the remainder goes to the low ranks and no sample is
    duplicated (unlike DistributedSampler's padding, which would change J)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)

"""I/O either side of the hot path.

Output side — restoration_test.py:133-157 calls ``torch.cuda.empty_cache()`` after every batch and then
``torchvision.utils.save_image`` once per image and kind (restored / low / sample / gt): each call moves one fp32 image to
the host, quantises it there and encodes a PNG on the main thread while the GPU idles.  Here

* ``quantize_u8`` does torchvision's normalise-and-quantise arithmetic on the DEVICE (``vsp_quantize_nchw_f32_to_hwc_u8``,
  byte-identical), so the device->host copy moves 3 bytes per pixel instead of 12, in one copy per batch;
* ``ImageWriter`` copies into pinned buffers on its own stream and encodes on a pool of host threads (OpenCV / PIL release
  the GIL), so encoding batch m overlaps the kernels of batch m+1; file names follow the reference
  (``{index:06d}_{rank}_{name}_{kind}.png``);
* nothing calls ``empty_cache`` (it forces a device synchronisation and hands the caching allocator's blocks back every batch).

Input side — the reference's test loader (dataset.py ``ImageFolder_restore_test``) decodes and resizes on the main process
with ``num_workers=0``; ``PrefetchLoader`` decodes on host threads into pinned batches and overlaps their host->device copy
with the previous batch's compute.  Training-time degradation (dataset.py:327-372: cv2 blur / resize / noise / JPEG) stays on
the CPU as in the reference — it needs a JPEG codec — but runs through the same prefetcher.
"""
from __future__ import annotations

import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from typing import Iterable, Iterator, List, Optional, Sequence

import numpy as np
import torch

from . import _lib


def quantize_u8(images: torch.Tensor, value_range=(-1.0, 1.0), out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[B,C,H,W] fp32 CUDA -> [B,H,W,C] uint8 CUDA with the bytes ``save_image(normalize=True, range=value_range)`` writes."""
    if not images.is_cuda:
        raise RuntimeError("vspbfr_b200.imageio.quantize_u8: CUDA tensor required (no CPU fallback)")
    x = images.contiguous().float()
    b, c, h, w = x.shape
    if out is None:
        out = torch.empty((b, h, w, c), dtype=torch.uint8, device=x.device)
    assert out.shape == (b, h, w, c) and out.is_contiguous() and out.dtype == torch.uint8
    with _lib.device_guard(x.device):
        rc = _lib.load().vsp_quantize_nchw_f32_to_hwc_u8(_lib.ptr(x), _lib.ptr(out), b, c, h * w, float(value_range[0]),
                                                         float(value_range[1]), _lib.stream_ptr())
    _lib.check(rc, "quantize_nchw_f32_to_hwc_u8")
    return out


def _encode_png(path: str, hwc_rgb: np.ndarray) -> None:
    try:
        import cv2

        arr = hwc_rgb if hwc_rgb.shape[2] == 1 else hwc_rgb[:, :, ::-1]      # RGB -> BGR view
        ok, buf = cv2.imencode(".png", np.ascontiguousarray(arr), [cv2.IMWRITE_PNG_COMPRESSION, 3])
        if not ok:
            raise RuntimeError("cv2.imencode failed")
        with open(path, "wb") as f:
            f.write(buf.tobytes())
    except ImportError:  # pragma: no cover
        from PIL import Image

        Image.fromarray(hwc_rgb.squeeze(-1) if hwc_rgb.shape[2] == 1 else hwc_rgb).save(path, format="PNG", compress_level=3)


class ImageWriter:
    """Asynchronous PNG writer for batches of device images.

        with ImageWriter(out_dir, rank=0, name="celeba") as wr:
            for i, batch in enumerate(loader):
                restored = ...
                wr.save(i * batch_size, restored=restored, low=low)     # returns immediately
    """

    def __init__(self, out_dir: str, rank: int = 0, name: str = "data", workers: int = 8, depth: int = 3):
        self.out_dir, self.rank, self.name = out_dir, rank, name
        os.makedirs(out_dir, exist_ok=True)
        self.pool = ThreadPoolExecutor(max_workers=workers)
        self._ring = [{"bufs": {}, "futs": []} for _ in range(max(1, depth))]
        self._k = 0
        self._stream = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def save(self, first_index: int, value_range=(-1.0, 1.0), **kinds: torch.Tensor) -> None:
        """Queue ``kinds`` (e.g. restored=, low=, sample=, gt=: [B,C,H,W] fp32 CUDA) for writing as
        ``{first_index + j:06d}_{rank}_{name}_{kind}.png`` (restoration_test.py:141-157)."""
        dev = next(iter(kinds.values())).device
        if self._stream is None:
            self._stream = torch.cuda.Stream(dev)
        slot = self._ring[self._k % len(self._ring)]
        self._k += 1
        for f in slot["futs"]:
            f.result()                                         # host threads that were still reading this slot's pinned buffers
        cur = torch.cuda.current_stream(dev)
        jobs = []
        for kind, img in kinds.items():
            b, c, h, w = img.shape
            bufs = slot["bufs"].get(kind)
            if bufs is None or tuple(bufs[0].shape) != (b, h, w, c):
                bufs = (torch.empty((b, h, w, c), dtype=torch.uint8, device=dev),
                        torch.empty((b, h, w, c), dtype=torch.uint8).pin_memory())
                slot["bufs"][kind] = bufs
            dbuf, hbuf = bufs
            quantize_u8(img, value_range, out=dbuf)            # on the compute stream, right behind the producer
            ev = torch.cuda.Event()
            ev.record(cur)
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(ev)
                hbuf.copy_(dbuf, non_blocking=True)
            jobs.append((kind, hbuf))
        done = torch.cuda.Event()
        done.record(self._stream)
        # the device buffers are reused `depth` batches later: by then this copy has long completed (checked via `done`
        # through the futures above)

        def encode(kind, hbuf, j):
            done.synchronize()
            _encode_png(os.path.join(self.out_dir, f"{first_index + j:06d}_{self.rank}_{self.name}_{kind}.png"), hbuf[j].numpy())

        slot["futs"] = [self.pool.submit(encode, kind, hbuf, j) for kind, hbuf in jobs for j in range(hbuf.shape[0])]

    def close(self) -> None:
        for slot in self._ring:
            for f in slot["futs"]:
                f.result()
            slot["futs"] = []
        self.pool.shutdown(wait=True)


def load_image(path: str, size: Optional[Sequence[int]] = None) -> np.ndarray:
    """Decode one image exactly as the reference's test datasets do (dataset.py:411-436, ImageFolder_restore_test
    ``__getitem__``): PIL RGB; if the size differs, LANCZOS-scale by max(H/h, W/w) and centre-crop to ``size`` = (H, W);
    then ToTensor + Normalize(0.5, 0.5) -> [-1, 1] fp32 CHW."""
    from PIL import Image

    img = Image.open(path).convert("RGB")
    if size is not None:
        w, h = img.size
        if h != size[0] or w != size[1]:
            ratio = max(1.0 * size[0] / h, 1.0 * size[1] / w)
            new_w, new_h = int(ratio * w), int(ratio * h)
            img = img.resize((new_w, new_h), Image.Resampling.LANCZOS)
            h_idx = (new_h - size[0]) // 2 if new_h - size[0] > 0 else 0
            w_idx = (new_w - size[1]) // 2 if new_w - size[1] > 0 else 0
            img = img.crop((w_idx, h_idx, int(w_idx + size[1]), int(h_idx + size[0])))
    x = np.asarray(img, dtype=np.uint8).astype(np.float32) / np.float32(255.0)      # ToTensor
    x = (x - np.float32(0.5)) / np.float32(0.5)                                      # Normalize((.5,.5,.5), (.5,.5,.5))
    return np.ascontiguousarray(x.transpose(2, 0, 1))


class PrefetchLoader:
    """Batches of decoded images as CUDA tensors, produced ahead of the consumer.

    ``paths`` -> iterator of (first_index, [B,3,H,W] fp32 CUDA).  Decoding runs on ``workers`` host threads, batches are
    assembled in pinned memory (``depth`` buffers) and copied on a side stream; the consumer's stream waits on the copy's
    event only, so host decode, H2D and the previous batch's kernels overlap."""

    def __init__(self, paths: Sequence[str], batch: int, size: Sequence[int], device="cuda", workers: int = 8, depth: int = 3,
                 decode=load_image):
        self.paths, self.batch, self.size = list(paths), int(batch), tuple(size)
        self.device = torch.device(device)
        self.workers, self.depth, self.decode = workers, depth, decode

    def __len__(self):
        return (len(self.paths) + self.batch - 1) // self.batch

    def __iter__(self) -> Iterator:
        n, bs = len(self.paths), self.batch
        pool = ThreadPoolExecutor(max_workers=self.workers)
        stream = torch.cuda.Stream(self.device)
        pinned = [torch.empty((bs, 3) + self.size, dtype=torch.float32).pin_memory() for _ in range(self.depth)]
        staged = [torch.empty((bs, 3) + self.size, dtype=torch.float32, device=self.device) for _ in range(self.depth)]
        free_q: "queue.Queue" = queue.Queue()                       # (buffer index, event: the consumer is done with it)
        for k in range(self.depth):
            free_q.put((k, None))
        ready_q: "queue.Queue" = queue.Queue()

        def produce():
            try:
                for s in range(0, n, bs):
                    k, ev_done = free_q.get()
                    if ev_done is not None:
                        ev_done.synchronize()                        # consumer kernels (and hence the old H2D copy) have finished
                    cnt = min(bs, n - s)
                    host = pinned[k]
                    list(pool.map(lambda j: host[j].copy_(torch.from_numpy(self.decode(self.paths[s + j], self.size))), range(cnt)))
                    with torch.cuda.stream(stream):
                        staged[k][:cnt].copy_(host[:cnt], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(stream)
                    ready_q.put((s, k, cnt, ev))
                ready_q.put(None)
            except BaseException as e:  # surface decode errors in the consumer
                ready_q.put(e)

        t = threading.Thread(target=produce, daemon=True)
        t.start()
        try:
            while True:
                item = ready_q.get()
                if item is None:
                    break
                if isinstance(item, BaseException):
                    raise item
                s, k, cnt, ev = item
                cur = torch.cuda.current_stream(self.device)
                cur.wait_event(ev)
                yield s, staged[k][:cnt]
                done = torch.cuda.Event()
                done.record(torch.cuda.current_stream(self.device))
                free_q.put((k, done))
        finally:
            pool.shutdown(wait=False)


def list_images(root: str) -> List[str]:
    """The files the reference's test datasets pick up (dataset.py:393-402: op.utils_train.listdir, sorted, by extension)."""
    from .op.utils_train import listdir

    return [p for p in listdir(root) if p.endswith((".JPG", ".jpg", ".png", ".jpeg"))]


def restore_folder(front, diffusion, decoder, net, lq_root: str, out_dir: str, batch: int = 4, size: int = 512,
                   name: str = "data", rank: int = 0, world: int = 1, device="cuda", workers: int = 8,
                   save_low: bool = True, save_sample: bool = True, pipeline=None) -> int:
    """restoration_test.py:111-157 (tester_restore_ddpm without ground truth) end to end: every image under ``lq_root`` is
    decoded ahead on host threads, restored (e4e encoder -> code diffusion -> style decoder + Restoration_net) and written
    as ``{index:06d}_{rank}_{name}_{restore|low|sample}.png`` by the asynchronous writer — decode, host->device copy,
    kernels, device->host copy and PNG encoding of neighbouring batches overlap; nothing synchronises per batch.
    ``pipeline`` (e.g. a ``frontend.GraphedPipeline`` captured for ``batch``) replaces the eager call for full batches.
    Ranks take contiguous slices (sharding.shard_range).  Returns the number of images this rank restored."""
    from . import frontend, sharding

    paths = list_images(lq_root)
    lo, hi = sharding.shard_range(len(paths), rank, world)
    dev = torch.device(device)
    done = 0
    with torch.no_grad(), ImageWriter(out_dir, rank=rank, name=name, workers=workers) as wr:
        for first, low in PrefetchLoader(paths[lo:hi], batch, (size, size), device=dev, workers=workers):
            z = torch.randn(low.shape[0], net.style_dim, device=dev)
            if pipeline is not None and low.shape[0] == getattr(pipeline, "micro", -1):
                restored, sample, _ = pipeline(low, z)
            else:
                restored, sample, _ = frontend.restore_pipeline(low, front, diffusion, decoder, net, [z])
            kinds = {"restore": restored}
            if save_low:
                kinds["low"] = low
            if save_sample:
                kinds["sample"] = sample
            wr.save(lo + first, **kinds)
            done += low.shape[0]
    return done

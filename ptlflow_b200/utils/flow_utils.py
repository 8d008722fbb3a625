"""Flow file I/O for the caller side of the hot path.

Same entry points and on-disk conventions as the reference's ``ptlflow/utils/flow_utils.py:60-162`` (which wraps
``utils/external/flowpy.py:218-352``) for the formats its ``infer.py`` / ``validate.py`` loops write:

* ``.flo``  Middlebury: ``b"PIEH"``, uint32 width, uint32 height, float32 [H,W,2]; NaN <-> the 1666666800.0 sentinel,
  anything above 1e9 in magnitude reads back as NaN
* ``.png``  KITTI: 16-bit RGB, R/G = flow * 64 + 2**15, B = valid (``png128``: multiplier 128)
* ``.npy``  plain ``numpy.save``

plus ``AsyncFlowWriter``: the reference writes every prediction synchronously between two forwards
(``infer.py:382-396``); here the D2H copy and the encode + write happen on a worker pool, off the critical path.
"""
from __future__ import annotations

import queue
import struct
import threading
from pathlib import Path
from typing import IO, Optional, Union

import numpy as np

_FLO_MAGIC = b"PIEH"
_FLO_SENTINEL = 1666666800.0
_FLO_INVALID_ABOVE = 1e9

PathLike = Union[str, Path]


def _format_of(path, format: Optional[str]) -> str:
    if format is not None:
        return format
    name = str(path)
    for ext in ("png128", "flo", "png", "npy"):
        if name.endswith(ext):
            return ext
    raise ValueError(f"cannot guess the flow format of {name!r}: pass format='flo' | 'png' | 'png128' | 'npy'")


# ---------------------------------------------------------------------------------------------
# .flo
# ---------------------------------------------------------------------------------------------
def _read_flo(f: IO[bytes]) -> np.ndarray:
    header = f.read(12)
    if len(header) != 12:
        raise ValueError("truncated .flo header")
    if header[:4] != _FLO_MAGIC:
        import warnings

        warnings.warn("file does not have a .flo signature")
    width, height = struct.unpack("<II", header[4:])
    data = np.frombuffer(f.read(width * height * 8), dtype="<f4")
    if data.size != width * height * 2:
        raise ValueError(f".flo payload has {data.size} values, header says {width}x{height}x2")
    flow = data.reshape(height, width, 2).astype(np.float32, copy=True)
    invalid = (np.abs(np.nan_to_num(flow, nan=0.0)) > _FLO_INVALID_ABOVE).any(axis=-1)
    flow[invalid] = np.nan
    return flow


def _write_flo(f: IO[bytes], flow: np.ndarray) -> None:
    height, width, _ = flow.shape
    payload = np.where(np.isnan(flow), np.float32(_FLO_SENTINEL), flow).astype("<f4")
    f.write(_FLO_MAGIC)
    f.write(struct.pack("<II", width, height))
    f.write(payload.tobytes())


# ---------------------------------------------------------------------------------------------
# KITTI .png
# ---------------------------------------------------------------------------------------------
def _read_png(path, mult: float) -> np.ndarray:
    import cv2

    if hasattr(path, "read"):  # file object: decode from its bytes
        bgr = cv2.imdecode(np.frombuffer(path.read(), dtype=np.uint8), cv2.IMREAD_UNCHANGED)
    else:
        bgr = cv2.imread(str(path), cv2.IMREAD_UNCHANGED)
    if bgr is None or bgr.ndim != 3 or bgr.shape[2] != 3 or bgr.dtype != np.uint16:
        raise ValueError(f"{path}: not a 16-bit 3-channel flow png")
    rgb = bgr[..., ::-1]
    flow = (rgb[..., :2].astype(np.float32) - 2**15) / mult
    flow[rgb[..., 2] == 0] = np.nan
    return flow


def _write_png(path, flow: np.ndarray, mult: float) -> None:
    import cv2

    valid = ~np.isnan(flow).any(axis=-1)
    uv = np.where(valid[..., None], flow, 0.0)
    enc = np.empty(flow.shape[:2] + (3,), dtype=np.uint16)
    enc[..., :2] = (uv * mult + 2**15).astype(np.uint16)
    enc[..., 2] = valid
    # always through the '.png' encoder (a '.png128' suffix has no cv2 writer); file objects get the encoded bytes
    ok, buf = cv2.imencode(".png", np.ascontiguousarray(enc[..., ::-1]))
    if not ok:
        raise OSError(f"could not encode {path} as png")
    if hasattr(path, "write"):
        path.write(buf.tobytes())
    else:
        with open(path, "wb") as f:
            f.write(buf.tobytes())


# ---------------------------------------------------------------------------------------------
# public entry points (names and argument meaning of ptlflow.utils.flow_utils)
# ---------------------------------------------------------------------------------------------
def flow_read(input_data: Union[PathLike, IO[bytes]], format: Optional[str] = None) -> np.ndarray:
    """Returns the flow as float32 [H,W,2] (x displacement first), invalid pixels NaN."""
    fmt = _format_of(getattr(input_data, "name", input_data), format)
    if fmt == "npy":
        return np.load(input_data)
    if fmt in ("png", "png128"):
        return _read_png(input_data, 128.0 if fmt == "png128" else 64.0)
    if fmt != "flo":
        raise ValueError(f"unsupported flow format {fmt!r}")
    if hasattr(input_data, "read"):
        return _read_flo(input_data)
    with open(input_data, "rb") as f:
        return _read_flo(f)


def flow_write(output_file: Union[PathLike, IO[bytes]], flow: np.ndarray, format: Optional[str] = None) -> None:
    """``flow``: [H,W,2], flow[..., 0] the x displacement; NaN marks invalid pixels."""
    flow = np.asarray(flow)
    if flow.ndim != 3 or flow.shape[2] != 2:
        raise ValueError(f"flow_write expects [H,W,2], got {flow.shape}")
    fmt = _format_of(getattr(output_file, "name", output_file), format)
    if fmt == "npy":
        np.save(output_file, flow)
    elif fmt in ("png", "png128"):
        _write_png(output_file, flow.astype(np.float32), 128.0 if fmt == "png128" else 64.0)
    elif fmt == "flo":
        if hasattr(output_file, "write"):
            _write_flo(output_file, flow.astype(np.float32))
        else:
            with open(output_file, "wb") as f:
                _write_flo(f, flow.astype(np.float32))
    else:
        raise ValueError(f"unsupported flow format {fmt!r}")


class AsyncFlowWriter:
    """Writes predicted flows on worker threads.

    ``submit(path, flow)`` takes a [2,H,W] or [H,W,2] tensor / array.  CUDA tensors are copied to pinned host memory on
    a side stream (ordered after the caller's current stream) and the caller returns immediately; encode + write run
    on the pool.  ``close()`` (or leaving the ``with`` block) waits for everything and re-raises the first error."""

    def __init__(self, workers: int = 2, max_pending: int = 32):
        self._q: "queue.Queue" = queue.Queue(maxsize=max_pending)
        self._error: Optional[BaseException] = None
        self._copy_stream = None
        self._threads = [threading.Thread(target=self._run, daemon=True, name=f"pfb-flow-writer{i}") for i in range(workers)]
        for t in self._threads:
            t.start()

    def _run(self) -> None:
        while True:
            job = self._q.get()
            try:
                if job is None:
                    return
                path, flow, event, fmt = job
                if event is not None:
                    event.synchronize()
                arr = flow.numpy() if hasattr(flow, "numpy") else np.asarray(flow)
                if arr.ndim == 3 and arr.shape[0] == 2 and arr.shape[2] != 2:
                    arr = arr.transpose(1, 2, 0)
                flow_write(path, arr.astype(np.float32), fmt)
            except BaseException as e:  # noqa: BLE001 -- reported by close()
                if self._error is None:
                    self._error = e
            finally:
                self._q.task_done()

    def submit(self, path: PathLike, flow, format: Optional[str] = None) -> None:
        event = None
        try:
            import torch

            if isinstance(flow, torch.Tensor):
                flow = flow.detach()
                if flow.is_cuda:
                    if self._copy_stream is None:
                        self._copy_stream = torch.cuda.Stream(device=flow.device)
                    host = torch.empty(flow.shape, dtype=torch.float32).pin_memory()
                    self._copy_stream.wait_stream(torch.cuda.current_stream(flow.device))
                    with torch.cuda.stream(self._copy_stream):
                        host.copy_(flow, non_blocking=True)
                        event = torch.cuda.Event()
                        event.record(self._copy_stream)
                    flow.record_stream(self._copy_stream)
                    flow = host
                else:
                    flow = flow.float()
        except ImportError:
            pass
        self._q.put((path, flow, event, format))

    def close(self) -> None:
        self._q.join()
        for _ in self._threads:
            self._q.put(None)
        for t in self._threads:
            t.join(timeout=60)
        if self._error is not None:
            raise self._error

    def __enter__(self) -> "AsyncFlowWriter":
        return self

    def __exit__(self, *exc) -> None:
        self.close()

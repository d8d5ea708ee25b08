"""Stream adapters over the batch API.

The reference's ``CompressionStream`` / ``DecompressionStream`` (src/ZstdSharp/CompressionStream.cs,
DecompressionStream.cs) drive zstd's streaming state machine, which is inherently serial.  These adapters keep the
classes' shape (constructor over an inner stream, ``Write``/``Flush``/``Read``/``Dispose``, ``SetParameter``) but cut the
data into independent frames of ``frame_size`` bytes and hand whole batches to ``Compressor.WrapBatch`` /
``Decompressor.UnwrapBatch``, so existing stream users reach the GPU path.  The output is format compatible:
concatenated frames are one valid zstd stream (``ZSTD_decompressMultiFrame``, Unsafe/ZstdDecompress.cs:1216), and
``DecompressionStream`` accepts any concatenation of frames (from this class, the reference, or the zstd CLI) and
decodes them as independent items after a host-side header walk (``ZSTD_findFrameCompressedSize``, :958).
"""
from __future__ import annotations

import io
from typing import List, Optional

import numpy as np

from . import _native
from .api import Compressor, Decompressor, EnsureZstdSuccess, ObjectDisposedException, ZstdException, ZSTD_ErrorCode, is_error, error_code

_lib = _native.lib

FRAME_SIZE = 128 * 1024          # one block per frame: the shape the batch kernels are built for
BATCH_FRAMES = 1024              # frames handed to the GPU at once (128 MiB of input)


class CompressionStream(io.RawIOBase):
    """Writable stream: everything written is compressed into independent frames and appended to ``stream``."""

    def __init__(self, stream, level: int = Compressor.DefaultCompressionLevel, bufferSize: int = 0,
                 compressor: Optional[Compressor] = None, frame_size: int = FRAME_SIZE, batch_frames: int = BATCH_FRAMES,
                 leaveOpen: bool = True):
        super().__init__()
        if stream is None:
            raise ValueError("stream")
        if not stream.writable():
            raise ValueError("Stream is not writable")
        if bufferSize < 0 or frame_size <= 0 or frame_size > FRAME_SIZE or batch_frames <= 0:
            raise ValueError("bufferSize / frame_size / batch_frames")
        self._inner = stream
        self._own = compressor is None
        self._comp = compressor if compressor is not None else Compressor(level)
        self._frame = frame_size
        self._batch = batch_frames
        self._leave = leaveOpen
        self._buf = bytearray()
        self._done = False

    # -- reference surface (CompressionStream.cs:46-75)
    def SetParameter(self, parameter: int, value: int) -> None:
        self._ensure()
        self._comp.SetParameter(parameter, value)

    def LoadDictionary(self, dictionary) -> None:     # CompressionStream.cs:58-62: every frame of the stream is compressed with it
        self._ensure()
        self._comp.LoadDictionary(dictionary)

    def writable(self) -> bool:
        return True

    def Write(self, buffer, offset: int = 0, count: Optional[int] = None) -> None:
        self._ensure()
        view = memoryview(buffer).cast("B")
        count = len(view) - offset if count is None else count
        self._buf += view[offset:offset + count]
        full = self._frame * self._batch
        while len(self._buf) >= full:
            self._emit(full)

    def write(self, b) -> int:            # io.RawIOBase protocol
        self.Write(b)
        return len(memoryview(b).cast("B"))

    def Flush(self) -> None:
        """Compresses everything buffered so far (the last frame may be short) and flushes the inner stream."""
        self._ensure()
        if self._buf:
            self._emit(len(self._buf))
        self._inner.flush()

    def flush(self) -> None:
        if not self._done:
            self.Flush()

    def _emit(self, nbytes: int) -> None:
        data = np.frombuffer(bytes(self._buf[:nbytes]), dtype=np.uint8)
        del self._buf[:nbytes]
        chunks = [data[i:i + self._frame] for i in range(0, data.size, self._frame)]
        for f in self._comp.WrapBatch(chunks):
            self._inner.write(f)

    def Dispose(self) -> None:
        if self._done:
            return
        try:
            self.Flush()
        finally:
            self._done = True
            if self._own:
                self._comp.Dispose()
            if not self._leave:
                self._inner.close()

    def close(self) -> None:
        self.Dispose()
        super().close()

    def _ensure(self) -> None:
        if self._done:
            raise ObjectDisposedException("CompressionStream")


def split_frames(blob: bytes) -> "tuple[List[bytes], int]":
    """Cuts a concatenation of zstd / skippable frames into its frames.  Returns (frames, consumed): a trailing
    incomplete frame is left unconsumed (more input may follow); garbage raises the error zstd would report."""
    a = np.frombuffer(blob, dtype=np.uint8)
    base = a.ctypes.data if a.size else 0
    frames, pos, n = [], 0, len(blob)
    while pos < n:
        r = int(_lib.ZSTD_findFrameCompressedSize(base + pos, n - pos))
        if is_error(r):
            if error_code(r) == ZSTD_ErrorCode.srcSize_wrong:
                break                              # incomplete frame: wait for more input
            EnsureZstdSuccess(r)
        frames.append(blob[pos:pos + r])
        pos += r
    return frames, pos


class DecompressionStream(io.RawIOBase):
    """Readable stream over a concatenation of zstd frames: frames are decoded in batches on the GPU."""

    def __init__(self, stream, bufferSize: int = 0, decompressor: Optional[Decompressor] = None,
                 checkEndOfStream: bool = True, leaveOpen: bool = True, batch_bytes: int = 64 << 20):
        super().__init__()
        if stream is None:
            raise ValueError("stream")
        if not stream.readable():
            raise ValueError("Stream is not readable")
        if bufferSize < 0:
            raise ValueError("bufferSize")
        self._inner = stream
        self._own = decompressor is None
        self._dec = decompressor if decompressor is not None else Decompressor()
        self._check = checkEndOfStream
        self._leave = leaveOpen
        self._batch_bytes = batch_bytes
        self._pending = b""           # compressed bytes not yet decoded
        self._out = bytearray()       # decoded bytes not yet handed out
        self._eof = False
        self._done = False

    def LoadDictionary(self, dictionary) -> None:     # DecompressionStream.cs:58-62
        self._dec.LoadDictionary(dictionary)

    def readable(self) -> bool:
        return True

    def _fill(self) -> None:
        while not self._out and not (self._eof and not self._pending):
            if not self._eof:
                more = self._inner.read(self._batch_bytes)
                if not more:
                    self._eof = True
                else:
                    self._pending += more
            frames, used = split_frames(self._pending)
            self._pending = self._pending[used:]
            regular = [f for f in frames if not (len(f) >= 4 and (int.from_bytes(f[:4], "little") & 0xFFFFFFF0) == 0x184D2A50)]
            if regular:
                for piece in self._dec.UnwrapBatch(regular):
                    self._out += piece
            if self._eof and self._pending:
                if self._check:         # DecompressionStream.cs: premature end of stream
                    raise ZstdException(ZSTD_ErrorCode.srcSize_wrong, "Premature end of stream")
                self._pending = b""

    def Read(self, count: int = -1) -> bytes:
        if self._done:
            raise ObjectDisposedException("DecompressionStream")
        if count is None or count < 0:
            chunks = []
            while True:
                self._fill()
                if not self._out:
                    break
                chunks.append(bytes(self._out)); self._out.clear()
            return b"".join(chunks)
        self._fill()
        piece = bytes(self._out[:count])
        del self._out[:count]
        return piece

    def readinto(self, b) -> int:        # io.RawIOBase protocol
        view = memoryview(b).cast("B")
        piece = self.Read(len(view))
        view[:len(piece)] = piece
        return len(piece)

    def readall(self) -> bytes:
        return self.Read(-1)

    def Dispose(self) -> None:
        if self._done:
            return
        self._done = True
        if self._own:
            self._dec.Dispose()
        if not self._leave:
            self._inner.close()

    def close(self) -> None:
        self.Dispose()
        super().close()

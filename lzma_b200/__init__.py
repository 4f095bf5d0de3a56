"""lzma_b200 -- B200-native batch LZMA / LZMA2 decoder behind the reader API of kulaginds/lzma.

Host-side mirror (Python over the C ABI in include/lzgpu.h; the reference is Go and the Go
toolchain is absent here -- the cgo binding a maintainer adds is in INTEGRATION.md and go/):

  reference (package lzma)                      here
  NewReader1 / (*Reader1).Read                  reader1.NewReader1 / Reader1.Read
  NewReader2 / (*Reader2).Read                  reader2.NewReader2 / Reader2.Read
  NewLZMADecompressorForSevenZip, NewLZMA2...   same names
  DecodeProp, DecodeDictSize, DecodeUnpackSize, DecodeDictSize2
  Err* (errors.go)                              errors.Err*, errors.Is
  (new) batch entry point                       batch.Context.decode_batch / lzgpu_decode_batch
  (new) .xz container front-end                 xz.decode_xz / xz.decode_xz_files (all blocks in one batch)

All decoding happens in liblzgpu.so on the GPU; importing the decode entry points without the
built library, or calling them without a CUDA device, fails loudly.
"""
from . import errors
from .errors import (EOF, ErrCorrupted, ErrDictOutOfRange, ErrIncorrectProperties, ErrNoLZMAReader, ErrResultError,
                     ErrUnexpectedEOF, ErrUnexpectedLZMA2Code, Is)


_LAZY = {
    "NewReader1": "reader1", "Reader1": "reader1", "NewLZMADecompressorForSevenZip": "reader1",
    "DecodeProp": "reader1", "DecodeDictSize": "reader1", "DecodeUnpackSize": "reader1",
    "NewReader2": "reader2", "Reader2": "reader2", "NewLZMA2DecompressorForSevenZip": "reader2",
    "DecodeDictSize2": "reader2", "Context": "batch", "Plan": "batch", "decode_alone_streams": "batch",
    "decode_lzma2_stream": "batch", "scan_lzma2": "batch", "shard_units": "batch",
    "decode_xz": "xz", "decode_xz_files": "xz", "scan_xz": "xz",
}


def __getattr__(name):
    # lazy: these pull in numpy and the ctypes binding of liblzgpu.so
    if name in _LAZY:
        import importlib
        return getattr(importlib.import_module(f"{__name__}.{_LAZY[name]}"), name)
    raise AttributeError(name)


def io_copy(dst, src, buf_size: int = 32 * 1024):
    """io.Copy: Read until EOF or error, writing to dst (an object with .write or .update)."""
    buf = bytearray(buf_size)
    total = 0
    write = getattr(dst, "write", None) or getattr(dst, "update")
    while True:
        n, err = src.Read(buf)
        if n:
            write(bytes(buf[:n]))
            total += n
        if err is not None:
            return total, (None if err is EOF else err)

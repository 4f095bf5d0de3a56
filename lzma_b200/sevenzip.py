"""All folders (coders) of a 7z archive in ONE GPU call.

bodgit/sevenzip hands each folder to the decompressor registered for its method -- `NewLZMADecompressorForSevenZip`
(reader1.go:32-61) for LZMA, `NewLZMA2DecompressorForSevenZip` (reader2.go:49-75) for LZMA2 -- one reader, and
with the GPU engine one call, per folder.  An archive reader that knows its folders up front decodes them
together: LZMA folders become headerless LZMA1 units, LZMA2 folders are cut at their dictionary resets, and the
whole archive is one `lzgpu_decode_batch`.  Mirror of `Engine::DecodeFolders` (lzma_reader.hpp) and of
`(*Engine).DecodeFolders` (go/lzgpu.go)."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _lib as L
from . import batch as B
from . import errors as E
from .reader1 import DecodeDictSize, DecodeProp, _status_error
from .reader2 import DecodeDictSize2


@dataclass
class Folder:
    lzma2: bool            # method 0x21 (props = 1 byte) rather than 0x030101 (props = 5 bytes)
    props: bytes
    unpack_size: int       # LZMA: from the archive header; LZMA2: ignored (the chunk headers say)
    packed: bytes


def decode_folders(ctx, folders: list[Folder]):
    """Returns [(bytes, err)] in folder order.  Property errors are the constructors' (ErrIncorrectProperties,
    errInsufficientProperties); decode errors are the readers', with the bytes decoded before them."""
    out: list = [None] * len(folders)
    caps: dict[int, int] = {}
    todo = []
    for i, f in enumerate(folders):
        if f.lzma2:
            if len(f.props) != 1:
                out[i] = (b"", E.errInsufficientProperties)
                continue
        else:
            if len(f.props) < 5 or DecodeProp(f.props[0])[3] is not None:
                out[i] = (b"", E.ErrIncorrectProperties)
                continue
            # the size field of an archive header is untrusted: first capacity bounded by the packed size
            caps[i] = min(f.unpack_size, max(1 << 16, 8 * len(f.packed)))
        todo.append(i)
    while todo:
        units, spans, chunks = [], {}, []
        in_off = out_off = 0
        for i in todo:
            f = folders[i]
            first = len(units)
            if f.lzma2:
                us, total, _sst = B.scan_lzma2(f.packed, DecodeDictSize2(f.props[0]))
                for u in us:
                    u.in_off += in_off
                    u.out_off += out_off
                units.extend(us)
                cap = total
            else:
                lc, pb, lp, _ = DecodeProp(f.props[0])
                u = L.Unit()
                u.kind = L.KIND_LZMA1_RAW
                u.lc, u.lp, u.pb = lc, lp, pb
                u.dict_size = DecodeDictSize(f.props[1:5])[0]
                u.unpack_size = f.unpack_size
                u.in_off, u.in_len, u.out_off, u.out_cap = in_off, len(f.packed), out_off, caps[i]
                units.append(u)
                cap = caps[i]
            spans[i] = (first, len(units) - first, out_off)
            chunks.append((in_off, f.packed))
            in_off = (in_off + len(f.packed) + 15) & ~15
            out_off = (out_off + cap + 15) & ~15
        in_buf = np.zeros(in_off + 16, dtype=np.uint8)
        for o, d in chunks:
            in_buf[o:o + len(d)] = np.frombuffer(d, dtype=np.uint8)
        out_buf = B._out_buffer(out_off + 16)
        res, _ = ctx.decode_batch(units, in_buf, out_buf)
        again = []
        for i in todo:
            first, cnt, o0 = spans[i]
            f = folders[i]
            if not f.lzma2 and res[first].status == L.OUTPUT_OVERFLOW and caps[i] < f.unpack_size:
                caps[i] = min(caps[i] * 8, f.unpack_size)     # the capacity guess was too small: this folder again
                again.append(i)
                continue
            n_out, err = 0, None
            for k in range(first, first + cnt):               # a folder's units are consecutive, and so are their outputs
                n_out = units[k].out_off - o0 + res[k].bytes_out
                if res[k].status != L.OK:
                    err = _status_error(res[k].status)
                    break
            out[i] = (out_buf[o0:o0 + n_out].tobytes(), err)
        todo = again
    return out

"""CPU reference of the PNG front end (TEST INFRASTRUCTURE like the rest of oracle/): scanline filters of the PNG
specification (section 9) with numpy, the adaptive choice (section 12.8: smallest sum of absolute values of the filtered
bytes read as signed, ties to the lower filter type), and a small decoder (chunk walk, CRC check, zlib inflate, unfilter)."""
import struct
import zlib

import numpy as np

CHANNELS = {0: 1, 2: 3, 4: 2, 6: 4}


def _neighbours(img, bpp):
    """img: (H, rowbytes) uint8 -> int32 arrays a (left), b (up), c (upper left) of the RAW samples."""
    h, w = img.shape
    raw = img.astype(np.int32)
    a = np.zeros_like(raw)
    a[:, bpp:] = raw[:, :-bpp] if w > bpp else 0
    b = np.zeros_like(raw)
    b[1:] = raw[:-1]
    c = np.zeros_like(raw)
    c[1:, bpp:] = raw[:-1, :-bpp] if w > bpp else 0
    return raw, a, b, c


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = np.abs(p - a), np.abs(p - b), np.abs(p - c)
    return np.where((pa <= pb) & (pa <= pc), a, np.where(pb <= pc, b, c))


def filter_image(pixels: bytes, width: int, height: int, color_type: int, filt: int) -> bytes:
    """-> height x (1 + rowbytes) filtered bytes; filt 0..4, or 5 = adaptive per scanline."""
    bpp = CHANNELS[color_type]
    img = np.frombuffer(pixels, dtype=np.uint8).reshape(height, width * bpp)
    raw, a, b, c = _neighbours(img, bpp)
    cand = [raw, raw - a, raw - b, raw - ((a + b) >> 1), raw - _paeth(a, b, c)]
    cand = [(x & 0xFF).astype(np.uint8) for x in cand]
    if filt < 5:
        types = np.full(height, filt, dtype=np.uint8)
    else:
        sums = np.stack([np.abs(x.view(np.int8).astype(np.int64)).sum(axis=1) for x in cand])   # (5, H)
        types = np.argmin(sums, axis=0).astype(np.uint8)                                       # first minimum
    out = np.empty((height, 1 + width * bpp), dtype=np.uint8)
    out[:, 0] = types
    for t in range(5):
        rows = types == t
        out[rows, 1:] = cand[t][rows]
    return out.tobytes()


def decode(png: bytes):
    """-> dict(width, height, color_type, pixels, idat=[payloads], zlib=concatenated IDAT payload).  Checks every CRC."""
    assert png[:8] == b"\x89PNG\r\n\x1a\n"
    pos, chunks = 8, []
    while pos < len(png):
        n, typ = struct.unpack(">I4s", png[pos:pos + 8])
        data = png[pos + 8:pos + 8 + n]
        crc, = struct.unpack(">I", png[pos + 8 + n:pos + 12 + n])
        assert zlib.crc32(typ + data) & 0xFFFFFFFF == crc, f"bad CRC in {typ!r}"
        chunks.append((typ, data))
        pos += 12 + n
    assert chunks[0][0] == b"IHDR" and chunks[-1] == (b"IEND", b"") and pos == len(png)
    w, h, depth, ct, comp, fm, il = struct.unpack(">IIBBBBB", chunks[0][1])
    assert (depth, comp, fm, il) == (8, 0, 0, 0)
    idat = [d for t, d in chunks if t == b"IDAT"]
    z = b"".join(idat)
    filt = zlib.decompress(z)
    bpp = CHANNELS[ct]
    rb = w * bpp
    assert len(filt) == h * (rb + 1)
    rows = np.frombuffer(filt, dtype=np.uint8).reshape(h, rb + 1)
    out = np.zeros((h, rb), dtype=np.uint8)
    prev = np.zeros(rb, dtype=np.int32)
    for y in range(h):
        t = int(rows[y, 0])
        line = rows[y, 1:].astype(np.int32)
        cur = np.zeros(rb, dtype=np.int32)
        if t == 0:
            cur = line
        elif t == 2:
            cur = (line + prev) & 0xFF
        else:   # Sub, Average, Paeth need the reconstructed left neighbour: one pass per pixel column
            for x in range(rb):
                a = cur[x - bpp] if x >= bpp else 0
                b = prev[x]
                c = prev[x - bpp] if x >= bpp else 0
                if t == 1:
                    pred = a
                elif t == 3:
                    pred = (a + b) >> 1
                else:
                    p = a + b - c
                    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[x] = (line[x] + pred) & 0xFF
        out[y] = cur
        prev = cur
    return dict(width=w, height=h, color_type=ct, pixels=out.tobytes(), idat=idat, zlib=z, filtered=filt)

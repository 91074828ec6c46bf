"""Test inputs: the Canterbury corpus fixture (tests/golden/canterbury.tar.xz, made by tests/golden/make_golden.py
from the reference's tests/corpus with the three text files restored to CRLF) and synthetic
generators for the BASELINE.json configs.  Never reads /root/reference."""
import io
import os
import tarfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURE = os.path.join(HERE, "golden", "canterbury.tar.xz")

_cache = None


def load_corpus():
    global _cache
    if _cache is None:
        out = {}
        with tarfile.open(FIXTURE, "r:xz") as tf:
            for m in tf.getmembers():
                out[m.name] = tf.extractfile(m).read()
        _cache = out
    return _cache


def sparse_image(width=3500, height=3500):
    """The reference's sparse RGBA image (DeflateSparseBenchmark.cs:83-99)."""
    y = np.arange(height, dtype=np.int64)[:, None]
    x4 = (np.arange(width, dtype=np.int64) * 4)[None, :]
    img = np.zeros((height, width, 4), dtype=np.uint8)
    img[:, :, 0] = ((x4 + y) % 256).astype(np.uint8)
    img[:, :, 3] = 255
    return img.tobytes()


_SYLL = ["th", "he", "in", "er", "an", "re", "on", "at", "en", "nd", "ti", "es", "or", "te", "of", "ed", "is", "it",
         "al", "ar", "st", "to", "nt", "ng", "se", "ha", "as", "ou", "io", "le", "ve", "co", "me", "de", "hi", "ri",
         "ro", "ic", "ne", "ea", "ra", "ce", "li", "ch", "ll", "be", "ma", "si", "om", "ur", "a", "e", "i", "o", "u",
         "s", "t", "y", "ly", "ing", "ion", "ent", "and", "the", "for", "ter", "ous", "ate", "ver", "all", "pro", "wh"]


def word_vocab(n_words=8192, seed=1):
    rng = np.random.RandomState(seed)
    words = []
    seen = set()
    while len(words) < n_words:
        k = 1 + min(4, int(rng.geometric(0.45)))
        w = "".join(_SYLL[i] for i in rng.randint(0, len(_SYLL), size=k))
        if w not in seen:
            seen.add(w)
            words.append(w)
    return words


def text_like(nbytes, seed=1):
    """Word-model English-like text (BASELINE config 2 "text-like"): Zipf-distributed words, sentence punctuation,
    line breaks about every 70 columns.  Deterministic in (nbytes, seed)."""
    rng = np.random.RandomState(seed)
    words = word_vocab()
    nw = len(words)
    p = 1.0 / np.arange(1, nw + 1) ** 1.05
    p /= p.sum()
    out = io.BytesIO()
    col = 0
    target = nbytes
    wb = [w.encode() for w in words]
    while out.tell() < target:
        idx = rng.choice(nw, size=4096, p=p)
        punct = rng.random_sample(4096)
        parts = []
        start = True
        for i, u in zip(idx, punct):
            w = wb[i]
            if start:
                w = w[:1].upper() + w[1:]
                start = False
            parts.append(w)
            col += len(w) + 1
            if u < 0.07:
                parts.append(b". ")
                start = True
            elif u < 0.15:
                parts.append(b", ")
            elif col > 68:
                parts.append(b"\r\n")
                col = 0
            else:
                parts.append(b" ")
        out.write(b"".join(parts))
    return out.getvalue()[:nbytes]


def mixed_entropy(nbytes, seed=2, tile=262144):
    """BASELINE config 3: tiles cycling text / record-like / sparse bitmap / random / zeros."""
    rng = np.random.RandomState(seed)
    out = bytearray()
    k = 0
    txt = text_like(4 * tile, seed + 100)
    while len(out) < nbytes:
        kind = k % 5
        if kind == 0:
            o = (k // 5 * tile) % (len(txt) - tile)
            out += txt[o:o + tile]
        elif kind == 1:  # spreadsheet-like fixed records with slowly varying fields
            n = tile // 16
            rec = np.zeros((n, 16), dtype=np.uint8)
            rec[:, 0] = 0x7E
            rec[:, 1:5] = (np.arange(n, dtype=np.uint32)[:, None] >> (8 * np.arange(4))[None, :]) & 0xFF
            rec[:, 5:9] = rng.randint(0, 4, size=(n, 4))
            rec[:, 9:13] = np.frombuffer(np.float32(rng.normal(100, 5, size=n)).round(1).tobytes(), dtype=np.uint8).reshape(n, 4)
            rec[:, 13:] = rng.randint(0, 256, size=(n, 3)) * (rng.random_sample((n, 1)) < 0.1)
            out += rec.tobytes()
        elif kind == 2:  # fax-like sparse bitmap: long zero runs with short bursts
            a = np.zeros(tile, dtype=np.uint8)
            pos = rng.randint(0, tile, size=tile // 64)
            a[pos] = rng.choice([0x01, 0x03, 0x80, 0xC0, 0xFF, 0x0F], size=pos.size)
            out += a.tobytes()
        elif kind == 3:
            out += rng.bytes(tile)
        else:
            out += bytes(tile)
        k += 1
    return bytes(out[:nbytes])


def png_like(nbytes, seed=3, width=4096):
    """BASELINE config 5: filtered RGBA scanlines (filter byte + Sub-filtered smooth gradient + noise)."""
    rng = np.random.RandomState(seed)
    row = 1 + 4 * width
    rows = nbytes // row + 1
    x = np.arange(width, dtype=np.float64)[None, :]
    y = np.arange(rows, dtype=np.float64)[:, None]
    img = np.zeros((rows, width, 4), dtype=np.int64)
    img[:, :, 0] = (127 + 120 * np.sin(x / 97.0 + y / 53.0)).astype(np.int64)
    img[:, :, 1] = (x / 16 + y / 8).astype(np.int64) % 256
    img[:, :, 2] = (127 + 100 * np.cos(x / 201.0) * np.sin(y / 77.0)).astype(np.int64)
    img[:, :, 3] = 255
    img[:, :, :3] += (rng.random_sample((rows, width, 3)) < 0.05) * rng.randint(-3, 4, size=(rows, width, 3))
    img &= 0xFF
    flat = img.reshape(rows, width * 4)
    sub = flat.copy()
    sub[:, 4:] = (flat[:, 4:] - flat[:, :-4]) & 0xFF   # PNG Sub filter
    out = np.empty((rows, row), dtype=np.uint8)
    out[:, 0] = 1
    out[:, 1:] = sub.astype(np.uint8)
    return out.tobytes()[:nbytes]

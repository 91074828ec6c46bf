"""the PNG front end on the device (zsb_png_encode).  Filtered scanlines must equal the numpy reference
(oracle/png_ref.py), the IDAT payload must equal the oracle's (= the reference deflater's) stream for those filtered bytes,
every chunk CRC must verify, and the decoded image must be the input."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))


def _image(w, h, ch, seed, kind="smooth"):
    rng = np.random.RandomState(seed)
    y, x = np.mgrid[0:h, 0:w]
    if kind == "noise":
        return rng.randint(0, 256, size=(h, w, ch)).astype(np.uint8).tobytes()
    base = (x * 3 + y * 2) % 256
    img = np.stack([(base * (k + 1) // 2 + rng.randint(0, 4, size=(h, w))) % 256 for k in range(ch)], axis=-1)
    return img.astype(np.uint8).tobytes()


@pytest.mark.parametrize("color_type", [0, 2, 4, 6])
@pytest.mark.parametrize("filt", [0, 1, 2, 3, 4, 5])
def test_png_filters_and_payload(engine, oracle, color_type, filt):
    import png_ref
    ch = png_ref.CHANNELS[color_type]
    w, h = 331, 97
    px = _image(w, h, ch, color_type * 10 + filt)
    png = engine.png_encode(px, w, h, color_type, filt, level=6, strategy=1)
    d = png_ref.decode(png)
    assert (d["width"], d["height"], d["color_type"]) == (w, h, color_type)
    assert d["pixels"] == px
    want_filtered = png_ref.filter_image(px, w, h, color_type, filt)
    assert d["filtered"] == want_filtered
    want_z, _, _ = oracle.deflate_segments(want_filtered, 65536, 6, 1)
    assert d["zlib"] == want_z          # the IDAT stream is the reference deflater's stream for these bytes


@pytest.mark.parametrize("level,strategy", [(6, 1), (6, 2), (6, 3), (1, 0), (9, 0), (0, 0)])
def test_png_rgba_4096_wide_every_strategy(engine, oracle, level, strategy):
    import png_ref
    w, h = 4096, 40           # 16385-byte scanlines (BASELINE config 5's shape), several 64 KiB segments
    px = _image(w, h, 4, 7)
    png = engine.png_encode(px, w, h, 6, 5, level=level, strategy=strategy)
    d = png_ref.decode(png)
    assert d["pixels"] == px
    filtered = png_ref.filter_image(px, w, h, 6, 5)
    assert d["filtered"] == filtered
    assert d["zlib"] == oracle.deflate_segments(filtered, 65536, level, strategy)[0]


def test_png_many_idat_chunks_and_their_crcs(engine):
    import png_ref
    w, h = 600, 500
    px = _image(w, h, 3, 3, kind="noise")      # incompressible: ~900 KB of zlib data
    png = engine.png_encode(px, w, h, 2, 4, level=1, strategy=0, idat_bytes=65536)
    d = png_ref.decode(png)                      # decode() checks the CRC of every chunk
    assert d["pixels"] == px
    assert len(d["idat"]) >= 10 and all(len(c) == 65536 for c in d["idat"][:-1]) and 0 < len(d["idat"][-1]) <= 65536
    one = engine.png_encode(px, w, h, 2, 4, level=1, strategy=0)
    assert len(png_ref.decode(one)["idat"]) == 1 and png_ref.decode(one)["zlib"] == d["zlib"]


def test_png_argument_errors(engine):
    with pytest.raises(Exception):
        engine.png_encode(b"\0" * 12, 2, 2, 3, 0)          # palette images are not built
    with pytest.raises(Exception, match="filter"):
        engine.png_encode(b"\0" * 16, 2, 2, 6, 9)

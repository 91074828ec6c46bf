"""Synthetic inputs for bench.py, generated with torch on the device the bench runs on (or on the CPU).

BASELINE.json config 2 "text-like": a word model -- Zipf-distributed tokens from a syllable vocabulary, with
punctuation and CRLF line ends as tokens of their own.  Level-1 ratio ~0.39-0.40, ~0.40 symbols per byte, ~0.28
searches per byte: the work profile of the Canterbury text files the reference's benchmarks use.
Deterministic in (nbytes, seed, chunk size); every 64 MiB chunk is generated independently.
"""
import numpy as np
import torch

_SYLL = ["th", "he", "in", "er", "an", "re", "on", "at", "en", "nd", "ti", "es", "or", "te", "of", "ed", "is", "it",
         "al", "ar", "st", "to", "nt", "ng", "se", "ha", "as", "ou", "io", "le", "ve", "co", "me", "de", "hi", "ri",
         "ro", "ic", "ne", "ea", "ra", "ce", "li", "ch", "ll", "be", "ma", "si", "om", "ur", "a", "e", "i", "o", "u",
         "s", "t", "y", "ly", "ing", "ion", "ent", "and", "the", "for", "ter", "ous", "ate", "ver", "all", "pro", "wh"]


def _vocab(n_words=8192, seed=1):
    rng = np.random.RandomState(seed)
    words, seen = [], set()
    while len(words) < n_words:
        k = 1 + min(4, int(rng.geometric(0.45)))
        w = "".join(_SYLL[i] for i in rng.randint(0, len(_SYLL), size=k))
        if w not in seen:
            seen.add(w)
            words.append(w)
    return words


def _token_table(device):
    words = _vocab()
    nw = len(words)
    p = 1.0 / np.arange(1, nw + 1) ** 1.05
    p /= p.sum()
    toks, probs = [], []
    for w, pw in zip(words, p):
        toks.append(w.encode() + b" ")
        probs.append(pw * 0.80)
        toks.append(w.encode() + b", ")
        probs.append(pw * 0.07)
        toks.append(w.encode() + b". ")
        probs.append(pw * 0.06)
        toks.append(w.encode() + b"\r\n")
        probs.append(pw * 0.07)
    flat = np.frombuffer(b"".join(toks), dtype=np.uint8).copy()
    lens = np.array([len(t) for t in toks], dtype=np.int64)
    starts = np.concatenate([[0], np.cumsum(lens)[:-1]])
    return (torch.from_numpy(flat).to(device), torch.from_numpy(starts).to(device), torch.from_numpy(lens).to(device),
            torch.tensor(probs, dtype=torch.float32, device=device), float(np.dot(lens, probs) / np.sum(probs)))


def text_like(nbytes: int, seed: int = 1, device="cuda", chunk: int = 64 << 20) -> torch.Tensor:
    """uint8 tensor of nbytes of word-model text on `device`."""
    flat, starts, lens, probs, mean_len = _token_table(device)
    out = torch.empty(nbytes, dtype=torch.uint8, device=device)
    pos, c = 0, 0
    while pos < nbytes:
        want = min(chunk, nbytes - pos)
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + c)
        ntok = int(want / mean_len * 1.05) + 64
        idx = torch.multinomial(probs, ntok, replacement=True, generator=g)
        l = lens[idx]
        ends = torch.cumsum(l, 0)
        j = torch.arange(want, device=device, dtype=torch.int64)
        t = torch.searchsorted(ends, j, right=True)
        t.clamp_(max=ntok - 1)
        tok_begin = ends[t] - l[t]
        off = (j - tok_begin).clamp_(min=0)
        src = starts[idx[t]] + torch.minimum(off, l[t] - 1)
        out[pos:pos + want] = flat[src]
        pos += want
        c += 1
        del idx, l, ends, j, t, tok_begin, off, src
    return out

"""Adler32 (Adler32.cs:19-78): `internal` in the reference, exposed here as the north star asks, plus Combine."""
from .engine import Engine, adler32_combine

_default_engine = None


def _engine() -> Engine:
    global _default_engine
    if _default_engine is None:
        _default_engine = Engine(0)
    return _default_engine


class Adler32:
    SeedValue = 1  # Adler32.cs:24

    @staticmethod
    def Calculate(buffer, adler: int = 1, engine: Engine = None) -> int:
        """Adler32.Calculate(adler, span) (Adler32.cs:61-78); an empty buffer returns the seed unchanged."""
        if len(buffer) == 0:
            return adler
        return (engine or _engine()).adler32(buffer, adler)

    @staticmethod
    def Combine(adler1: int, adler2: int, len2: int) -> int:
        return adler32_combine(adler1, adler2, len2)

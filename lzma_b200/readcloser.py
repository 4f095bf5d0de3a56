"""readcloser.go: pairs the decoded reader with the source's Close and wraps errors."""
from __future__ import annotations

from . import errors as E


class readCloser:
    def __init__(self, c, r):
        self.c, self.r = c, r

    def Close(self):
        if self.c is None or self.r is None:
            return E.errAlreadyClosed                         # readcloser.go:17-19
        try:
            self.c.close()
        except Exception as ex:  # the source's Close failed
            return E.Errorf("lzma: error closing", E.Error(str(ex)))
        self.c = self.r = None
        return None

    def Read(self, p):
        if self.r is None:
            return 0, E.errAlreadyClosed
        n, err = self.r.Read(p)
        if err is not None and not E.Is(err, E.EOF):
            err = E.Errorf("lzma: error reading", err)        # readcloser.go:36-38
        return n, err

"""Error values of the reference package (errors.go:5-12) and the Go std-lib sentinels it
returns, as Python objects.  Functions in this package return ``(value, err)`` pairs like
their Go originals; ``err`` is None or an :class:`Error`; use :func:`Is` like ``errors.Is``."""
from __future__ import annotations


class Error(Exception):
    """An error value.  ``wrapped`` mirrors fmt.Errorf("...: %w", err)."""

    def __init__(self, msg: str, wrapped: "Error | None" = None):
        super().__init__(msg)
        self.msg = msg
        self.wrapped = wrapped

    def __str__(self) -> str:
        return self.msg

    def __repr__(self) -> str:
        return f"Error({self.msg!r})"


def Errorf(prefix: str, err: Error) -> Error:
    """fmt.Errorf(prefix + ": %w", err)"""
    return Error(f"{prefix}: {err.msg}", wrapped=err)


def Is(err, target) -> bool:
    """errors.Is"""
    while err is not None:
        if err is target:
            return True
        err = getattr(err, "wrapped", None)
    return False


# io
EOF = Error("EOF")
ErrUnexpectedEOF = Error("unexpected EOF")

# errors.go:5-12 (ErrCorrupted, ErrDictOutOfRange, ErrUnexpectedLZMA2Code and ErrNoLZMAReader are
# never returned by the reference either; they exist so that callers' errors.Is checks keep compiling)
ErrCorrupted = Error("corrupted")
ErrIncorrectProperties = Error("incorrect LZMA properties")
ErrResultError = Error("result error")
ErrDictOutOfRange = Error("dictionary capacity is out of range")
ErrUnexpectedLZMA2Code = Error("unexpected lzma2 code")
ErrNoLZMAReader = Error("no lzma reader on chunkLZMAResetState")

# unexported in the reference (reader1.go:26, reader2.go:43, readcloser.go:14)
errNeedOneReader = Error("lzma: need exactly one reader")
errInsufficientProperties = Error("lzma2: not enough properties")
errAlreadyClosed = Error("lzma: already closed")

# new: the batch API needs a buffer size; the streaming readers never surface this
ErrOutputOverflow = Error("lzgpu: output capacity too small")

"""Comparison rules between the oracle and a decode backend.

Bit-exact bytes whenever the reference delivers them (status OK / OK_INPUT_EXHAUSTED);
on an error the reference delivers a caller-buffer-dependent prefix (reader1.go:245-252),
so parity there is the error class and, for LZMA1, the decompress.go site."""
from oracle import oracle as O


def same_outcome(want, got_status, got_site, got_data, name="", strict_site=True):
    assert got_status == want.status, f"{name}: status {got_status} != oracle {want.status_name} (site {want.err_site}/{got_site})"
    if want.status in (O.OK, O.OK_INPUT_EXHAUSTED):
        assert len(got_data) == len(want.data), f"{name}: {len(got_data)} bytes != oracle {len(want.data)}"
        assert got_data == want.data, f"{name}: decoded bytes differ"
    elif want.status == O.RESULT_ERROR and strict_site:
        assert got_site == want.err_site, f"{name}: error site {got_site} != oracle {want.err_site}"

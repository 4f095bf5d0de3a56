// Package lzma: cgo binding of liblzgpu.so (include/lzgpu.h) for kulaginds/lzma.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain. This file is the
// reference-side binding a maintainer adds next to reader1.go / reader2.go; the same C ABI is
// exercised from Python (lzma_b200/_lib.py) by the test-suite.
//
// Build:  CGO_CFLAGS="-I${LZGPU}/include" CGO_LDFLAGS="-L${LZGPU}/lzma_b200 -llzgpu" go build
package lzma

/*
#cgo LDFLAGS: -llzgpu
#include <stdlib.h>
#include "lzgpu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"io"
	"sync"
	"unsafe"
)

// ErrOutputOverflow has no analogue in the streaming readers: the batch API needs a capacity.
var ErrOutputOverflow = errors.New("lzgpu: output capacity too small")

// Unit describes one independently decodable piece of a batch (C.lzgpu_unit).
type Unit struct {
	Kind       uint8 // C.LZGPU_KIND_*
	In         []byte
	OutCap     uint64
	Lc, Lp, Pb uint8  // LZMA1_RAW only
	DictSize   uint32 // LZMA1_RAW / LZMA2
	UnpackSize uint64 // LZMA1_RAW; math.MaxUint64 = unknown
	Flags      uint32
}

// Result is the outcome of one unit (C.lzgpu_result) plus its decoded bytes.
type Result struct {
	Err     error // nil, io.EOF-equivalent statuses are mapped to nil
	Site    int   // decompress.go line of the failing return (diagnostic)
	Out     []byte
	BytesIn uint64
}

// Engine owns a C.lzgpu_ctx (streams + staging buffers on every visible GPU).
type Engine struct {
	mu  sync.Mutex
	ctx *C.lzgpu_ctx
}

// NewEngine fails when no CUDA device is present: there is no CPU fallback.
func NewEngine() (*Engine, error) {
	e := &Engine{}
	if rc := C.lzgpu_ctx_create(nil, 0, &e.ctx); rc != C.LZGPU_E_OK {
		return nil, fmt.Errorf("lzgpu: %s", C.GoString(C.lzgpu_last_error()))
	}
	return e, nil
}

func (e *Engine) Close() { C.lzgpu_ctx_destroy(e.ctx); e.ctx = nil }

// PinnedBuffer is host memory the GPUs can address (lzgpu_alloc_pinned): input placed in one is read by the
// decode kernel directly over PCIe, output written to one is streamed back while the kernel runs.
// DecodeBatch above lays its units into ordinary Go slices (staged through device slabs by the library); a caller
// with long-lived buffers holds PinnedBuffers and passes their .B to lzgpu_decode_batch for the zero-copy path.
type PinnedBuffer struct {
	p unsafe.Pointer
	B []byte
}

func NewPinnedBuffer(size int) (*PinnedBuffer, error) {
	p := C.lzgpu_alloc_pinned(C.uint64_t(size))
	if p == nil {
		return nil, fmt.Errorf("lzgpu: %s", C.GoString(C.lzgpu_last_error()))
	}
	return &PinnedBuffer{p: p, B: unsafe.Slice((*byte)(p), size)}, nil
}

func (b *PinnedBuffer) Free() { C.lzgpu_free_pinned(b.p); b.p, b.B = nil, nil }

// statusErr maps a per-unit status onto the package's error values (errors.go:5-12).
func statusErr(st C.int32_t) error {
	switch st {
	case C.LZGPU_OK, C.LZGPU_OK_INPUT_EXHAUSTED: // reader1.go:246-249 treats exhaustion as EOF
		return nil
	case C.LZGPU_RESULT_ERROR:
		return ErrResultError
	case C.LZGPU_INCORRECT_PROPERTIES:
		return ErrIncorrectProperties
	case C.LZGPU_UNEXPECTED_EOF:
		return io.ErrUnexpectedEOF
	case C.LZGPU_OUTPUT_OVERFLOW:
		return ErrOutputOverflow
	}
	return fmt.Errorf("lzgpu: status %d", int(st))
}

// DecodeBatch is the batch entry point: one goroutine-safe, synchronous call decodes all units,
// sharded over the engine's GPUs by compressed size.
func (e *Engine) DecodeBatch(units []Unit) ([]Result, error) {
	n := len(units)
	if n == 0 {
		return nil, nil
	}
	// lay the inputs into one buffer (16-byte aligned slots) and size the output
	cu := make([]C.lzgpu_unit, n)
	var inSize, outSize uint64
	for i := range units {
		cu[i].in_off, cu[i].in_len = C.uint64_t(inSize), C.uint64_t(len(units[i].In))
		cu[i].out_off, cu[i].out_cap = C.uint64_t(outSize), C.uint64_t(units[i].OutCap)
		cu[i].kind, cu[i].flags = C.uint8_t(units[i].Kind), C.uint32_t(units[i].Flags)
		cu[i].lc, cu[i].lp, cu[i].pb = C.uint8_t(units[i].Lc), C.uint8_t(units[i].Lp), C.uint8_t(units[i].Pb)
		cu[i].dict_size, cu[i].unpack_size = C.uint32_t(units[i].DictSize), C.uint64_t(units[i].UnpackSize)
		inSize = (inSize + uint64(len(units[i].In)) + 15) &^ 15
		outSize = (outSize + units[i].OutCap + 15) &^ 15
	}
	in := make([]byte, inSize+16)
	out := make([]byte, outSize+16)
	for i := range units {
		copy(in[cu[i].in_off:], units[i].In)
	}
	res := make([]C.lzgpu_result, n)
	e.mu.Lock()
	rc := C.lzgpu_decode_batch(e.ctx, &cu[0], C.int64_t(n),
		(*C.uint8_t)(unsafe.Pointer(&in[0])), C.uint64_t(len(in)),
		(*C.uint8_t)(unsafe.Pointer(&out[0])), C.uint64_t(len(out)), &res[0], nil)
	e.mu.Unlock()
	if rc != C.LZGPU_E_OK {
		return nil, fmt.Errorf("lzgpu: %s", C.GoString(C.lzgpu_last_error()))
	}
	r := make([]Result, n)
	for i := range r {
		o := uint64(cu[i].out_off)
		r[i] = Result{Err: statusErr(res[i].status), Site: int(res[i].err_site),
			Out: out[o : o+uint64(res[i].bytes_out)], BytesIn: uint64(res[i].bytes_in)}
	}
	return r, nil
}

// ScanLZMA2 is the host chunk scanner (Reader2.startChunk's framing rules, reader2.go:100-214).
func ScanLZMA2(data []byte, dictSize uint32) (units []C.lzgpu_unit, total uint64, truncated bool) {
	var tot C.uint64_t
	var sst C.int32_t
	var p *C.uint8_t
	if len(data) > 0 {
		p = (*C.uint8_t)(unsafe.Pointer(&data[0]))
	}
	n := C.lzgpu_scan_lzma2(p, C.uint64_t(len(data)), C.uint32_t(dictSize), nil, 0, &tot, &sst)
	units = make([]C.lzgpu_unit, n)
	if n > 0 {
		C.lzgpu_scan_lzma2(p, C.uint64_t(len(data)), C.uint32_t(dictSize), &units[0], n, &tot, &sst)
	}
	return units, uint64(tot), sst == C.LZGPU_UNEXPECTED_EOF
}

// Package lzma: cgo binding of liblzgpu.so (include/lzgpu.h) for kulaginds/lzma.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain. This file is the
// reference-side binding a maintainer adds next to reader1.go / reader2.go. Its marshalling of
// Unit -> C.lzgpu_unit is restated field by field in C (tests/cpp/go_marshal_test.c) and run against
// the GPU by the test-suite, so that the layout this file relies on cannot rot unnoticed; the same C
// ABI is also exercised from Python (lzma_b200/_lib.py) and C++ (lzma_b200/csrc/lzma_reader.cpp).
//
// Build:  CGO_CFLAGS="-I${LZGPU}/include" CGO_LDFLAGS="-L${LZGPU}/lzma_b200 -llzgpu" go build
package lzma

/*
#cgo LDFLAGS: -llzgpu
#include <stdlib.h>
#include <string.h>
#include "lzgpu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"io"
	"runtime"
	"sync"
	"unsafe"
)

// ErrOutputOverflow has no analogue in the streaming readers: the batch API needs a capacity.
var ErrOutputOverflow = errors.New("lzgpu: output capacity too small")

// Unit kinds (C.LZGPU_KIND_*).
const (
	KindLZMA1Alone = 0 // 13-byte .lzma header in the stream (NewReader1, reader1.go:18)
	KindLZMA1Raw   = 1 // lc/lp/pb/dict/size supplied (NewLZMADecompressorForSevenZip, reader1.go:32)
	KindLZMA2Group = 2 // LZMA2 chunk run starting at a dictionary reset (NewReader2, reader2.go:26)
)

// Unit describes one independently decodable piece of a batch (C.lzgpu_unit).
type Unit struct {
	Kind       uint8 // Kind*
	In         []byte
	OutCap     uint64
	Lc, Lp, Pb uint8  // LZMA1_RAW: the stream's; LZMA2: properties in force before the unit
	DictSize   uint32 // LZMA1_RAW / LZMA2
	UnpackSize uint64 // LZMA1_RAW; math.MaxUint64 = unknown
	Flags      uint32 // C.LZGPU_UF_*
	// LZMA2 only: the table sizes ScanLZMA2 computed (valid when Flags has LZGPU_UF_BITS_KNOWN). A Unit
	// built by hand leaves them 0 and the flag clear; the library then walks the chunk headers itself.
	LitBits, PosBits uint8
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

// lastError fetches the library's error text. lzgpu_last_error() is thread-local in the library, so the
// failing call and this one must run on the same OS thread: callers hold runtime.LockOSThread().
func lastError() error { return fmt.Errorf("lzgpu: %s", C.GoString(C.lzgpu_last_error())) }

// NewEngine fails when no CUDA device is present: there is no CPU fallback.
func NewEngine() (*Engine, error) {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	e := &Engine{}
	if rc := C.lzgpu_ctx_create(nil, 0, &e.ctx); rc != C.LZGPU_E_OK {
		return nil, lastError()
	}
	return e, nil
}

func (e *Engine) Close() { C.lzgpu_ctx_destroy(e.ctx); e.ctx = nil }

// PinnedBuffer is host memory the GPUs can address (lzgpu_alloc_pinned): input placed in one is read by the
// decode kernel directly over PCIe, output written to one is written there by the decoding units themselves, block by block, while the kernel runs.
type PinnedBuffer struct {
	p unsafe.Pointer
	B []byte
}

func NewPinnedBuffer(size int) (*PinnedBuffer, error) {
	runtime.LockOSThread()
	defer runtime.UnlockOSThread()
	if size < 1 {
		size = 1
	}
	p := C.lzgpu_alloc_pinned(C.uint64_t(size))
	if p == nil {
		return nil, lastError()
	}
	return &PinnedBuffer{p: p, B: unsafe.Slice((*byte)(p), size)}, nil
}

func (b *PinnedBuffer) Free() {
	if b != nil && b.p != nil {
		C.lzgpu_free_pinned(b.p)
		b.p, b.B = nil, nil
	}
}

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

// marshal fills C.lzgpu_unit i from units[i]; in_off / out_off are the 16-byte aligned slots the
// caller laid out. (tests/cpp/go_marshal_test.c restates exactly these assignments.)
func marshal(cu *C.lzgpu_unit, u *Unit, inOff, outOff uint64) {
	cu.in_off, cu.in_len = C.uint64_t(inOff), C.uint64_t(len(u.In))
	cu.out_off, cu.out_cap = C.uint64_t(outOff), C.uint64_t(u.OutCap)
	cu.kind, cu.flags = C.uint8_t(u.Kind), C.uint32_t(u.Flags)
	cu.lc, cu.lp, cu.pb = C.uint8_t(u.Lc), C.uint8_t(u.Lp), C.uint8_t(u.Pb)
	cu.lit_bits, cu.pos_bits = C.uint8_t(u.LitBits), C.uint8_t(u.PosBits)
	cu.dict_size, cu.unpack_size = C.uint32_t(u.DictSize), C.uint64_t(u.UnpackSize)
}

// DecodeBatch is the batch entry point: one goroutine-safe, synchronous call decodes all units,
// sharded over the engine's GPUs by compressed size. Inputs and outputs are ordinary Go slices (the
// library stages them through device slabs); DecodeBatchPinned is the zero-copy variant.
func (e *Engine) DecodeBatch(units []Unit) ([]Result, error) {
	n := len(units)
	if n == 0 {
		return nil, nil
	}
	cu, inSize, outSize := layout(units)
	in := make([]byte, inSize+16)
	out := make([]byte, outSize+16)
	for i := range units {
		copy(in[cu[i].in_off:], units[i].In)
	}
	res, err := e.call(cu, in, out)
	if err != nil {
		return nil, err
	}
	return results(cu, res, out), nil
}

// DecodeBatchPinned is DecodeBatch with caller-held page-locked buffers: the compressed inputs are laid into
// in.B (read by the kernel straight from host memory) and the decoded bytes land in out.B (written by the units themselves while
// the kernel runs). The Results' Out slices alias out.B and are valid until the caller reuses or frees it.
// ErrOutputOverflow-style sizing is the caller's: len(in.B) / len(out.B) must cover the laid-out units.
func (e *Engine) DecodeBatchPinned(units []Unit, in, out *PinnedBuffer) ([]Result, error) {
	n := len(units)
	if n == 0 {
		return nil, nil
	}
	cu, inSize, outSize := layout(units)
	if uint64(len(in.B)) < inSize || uint64(len(out.B)) < outSize {
		return nil, fmt.Errorf("lzgpu: pinned buffers too small (need %d in, %d out)", inSize, outSize)
	}
	for i := range units {
		copy(in.B[cu[i].in_off:], units[i].In)
	}
	res, err := e.call(cu, in.B, out.B)
	if err != nil {
		return nil, err
	}
	return results(cu, res, out.B), nil
}

// layout assigns 16-byte aligned input and output slots.
func layout(units []Unit) (cu []C.lzgpu_unit, inSize, outSize uint64) {
	cu = make([]C.lzgpu_unit, len(units)) // zeroed: pad8 / user stay 0
	for i := range units {
		marshal(&cu[i], &units[i], inSize, outSize)
		inSize = (inSize + uint64(len(units[i].In)) + 15) &^ 15
		outSize = (outSize + units[i].OutCap + 15) &^ 15
	}
	return cu, inSize, outSize
}

func (e *Engine) call(cu []C.lzgpu_unit, in, out []byte) ([]C.lzgpu_result, error) {
	res := make([]C.lzgpu_result, len(cu))
	var pin, pout *C.uint8_t
	if len(in) > 0 {
		pin = (*C.uint8_t)(unsafe.Pointer(&in[0]))
	}
	if len(out) > 0 {
		pout = (*C.uint8_t)(unsafe.Pointer(&out[0]))
	}
	runtime.LockOSThread() // the error text is thread-local in the library
	defer runtime.UnlockOSThread()
	e.mu.Lock()
	rc := C.lzgpu_decode_batch(e.ctx, &cu[0], C.int64_t(len(cu)), pin, C.uint64_t(len(in)), pout, C.uint64_t(len(out)), &res[0], nil)
	e.mu.Unlock()
	if rc != C.LZGPU_E_OK {
		return nil, lastError()
	}
	return res, nil
}

func results(cu []C.lzgpu_unit, res []C.lzgpu_result, out []byte) []Result {
	r := make([]Result, len(cu))
	for i := range r {
		o := uint64(cu[i].out_off)
		r[i] = Result{Err: statusErr(res[i].status), Site: int(res[i].err_site),
			Out: out[o : o+uint64(res[i].bytes_out)], BytesIn: uint64(res[i].bytes_in)}
	}
	return r
}

// ScanLZMA2 is the host chunk scanner (Reader2.startChunk's framing rules, reader2.go:100-214): the units of
// one raw LZMA2 stream, ready for DecodeBatch (In aliases data; OutCap is what the chunk headers promise).
func ScanLZMA2(data []byte, dictSize uint32) (units []Unit, total uint64, truncated bool) {
	var tot C.uint64_t
	var sst C.int32_t
	var p *C.uint8_t
	if len(data) > 0 {
		p = (*C.uint8_t)(unsafe.Pointer(&data[0]))
	}
	n := C.lzgpu_scan_lzma2(p, C.uint64_t(len(data)), C.uint32_t(dictSize), nil, 0, &tot, &sst)
	if n <= 0 {
		return nil, uint64(tot), sst == C.LZGPU_UNEXPECTED_EOF
	}
	cu := make([]C.lzgpu_unit, n)
	C.lzgpu_scan_lzma2(p, C.uint64_t(len(data)), C.uint32_t(dictSize), &cu[0], n, &tot, &sst)
	units = make([]Unit, n)
	for i := range cu {
		u := &cu[i]
		units[i] = Unit{Kind: KindLZMA2Group, In: data[u.in_off : u.in_off+u.in_len], OutCap: uint64(u.out_cap),
			Lc: uint8(u.lc), Lp: uint8(u.lp), Pb: uint8(u.pb), DictSize: uint32(u.dict_size),
			UnpackSize: uint64(u.unpack_size), Flags: uint32(u.flags),
			LitBits: uint8(u.lit_bits), PosBits: uint8(u.pos_bits)}
	}
	return units, uint64(tot), sst == C.LZGPU_UNEXPECTED_EOF
}

// Folder is one coder of a 7z archive as bodgit/sevenzip hands it to a registered decompressor
// (reader1.go:32-61, reader2.go:49-75): the coder's property bytes, the unpacked size, the packed bytes.
type Folder struct {
	LZMA2      bool   // method 0x21 (props = 1 byte) rather than 0x030101 (props = 5 bytes)
	Props      []byte
	UnpackSize uint64
	Packed     []byte
}

// DecodeFolders decodes all folders of an archive with ONE GPU call (instead of one reader, and one GPU
// call, per folder): LZMA folders become LZMA1_RAW units, LZMA2 folders are scanned into their units.
// Returns each folder's bytes and error in order; the property / argument errors are the constructors'
// (ErrIncorrectProperties, errInsufficientProperties).
func (e *Engine) DecodeFolders(folders []Folder) ([][]byte, []error, error) {
	outs := make([][]byte, len(folders))
	errs := make([]error, len(folders))
	var units []Unit
	var owner []int
	for i, f := range folders {
		if f.LZMA2 {
			if len(f.Props) != 1 {
				errs[i] = errInsufficientProperties
				continue
			}
			us, _, _ := ScanLZMA2(f.Packed, DecodeDictSize2(f.Props[0]))
			for range us {
				owner = append(owner, i)
			}
			units = append(units, us...)
			continue
		}
		if len(f.Props) < 5 {
			errs[i] = ErrIncorrectProperties
			continue
		}
		lc, pb, lp, err := DecodeProp(f.Props[0])
		if err != nil {
			errs[i] = err
			continue
		}
		ds, _ := DecodeDictSize(f.Props[1:5])
		units = append(units, Unit{Kind: KindLZMA1Raw, In: f.Packed, OutCap: f.UnpackSize, Lc: lc, Lp: lp, Pb: pb,
			DictSize: ds, UnpackSize: f.UnpackSize})
		owner = append(owner, i)
	}
	res, err := e.DecodeBatch(units)
	if err != nil {
		return nil, nil, err
	}
	for k, r := range res { // a folder's units are consecutive: deliver up to the first failing one
		i := owner[k]
		if errs[i] != nil {
			continue
		}
		outs[i] = append(outs[i], r.Out...)
		errs[i] = r.Err
	}
	return outs, errs, nil
}

// Drop-in replacements for the bodies of Reader1.Read / Reader2.Read (reader1.go:223-254,
// reader2.go:216-250): same exported API and error values, the in-stream decode runs on the GPU.
//
// NOT COMPILED HERE (no Go toolchain in the build image); see INTEGRATION.md.
package lzma

import (
	"bufio"
	"io"
	"math"
)

var defaultEngine *Engine

func engine() (*Engine, error) {
	if defaultEngine == nil {
		e, err := NewEngine()
		if err != nil {
			return nil, err // no device: no fallback, the caller sees the error
		}
		defaultEngine = e
	}
	return defaultEngine, nil
}

// gpuReader1 keeps the constructor behaviour of NewReader1 (header + range-coder preamble are read
// and validated eagerly, reader1.go:77-159) and decodes the body on the first Read.
type gpuReader1 struct {
	in                 io.ByteReader
	lc, lp, pb         uint8
	dictSize           uint32
	unpackSize         uint64
	preamble           [5]byte
	out                []byte
	pos                int
	err                error
	decoded, endOfData bool
}

func (r *gpuReader1) decode() {
	payload := append([]byte{}, r.preamble[:]...)
	for {
		b, err := r.in.ReadByte()
		if err != nil {
			break
		}
		payload = append(payload, b)
	}
	e, err := engine()
	if err != nil {
		r.err = err
		return
	}
	capacity := r.unpackSize
	if capacity == math.MaxUint64 {
		capacity = uint64(len(payload))*8 + 1<<16
	}
	for {
		res, err := e.DecodeBatch([]Unit{{Kind: 1, In: payload, OutCap: capacity, Lc: r.lc, Lp: r.lp, Pb: r.pb,
			DictSize: r.dictSize, UnpackSize: r.unpackSize}})
		if err != nil {
			r.err = err
			return
		}
		if res[0].Err == ErrOutputOverflow && r.unpackSize == math.MaxUint64 {
			capacity *= 8 // the streaming API has no capacity: grow and decode again
			continue
		}
		r.out, r.err = res[0].Out, res[0].Err
		return
	}
}

func (r *gpuReader1) Read(p []byte) (int, error) {
	if !r.decoded {
		r.decode()
		r.decoded = true
	}
	n := copy(p, r.out[r.pos:])
	r.pos += n
	if n == len(p) && n > 0 {
		return n, nil
	}
	if r.err != nil {
		err := r.err
		r.err = nil
		return n, err
	}
	return n, io.EOF
}

// gpuReader2: NewReader2's eager first-header read stays in the constructor; the first Read scans
// the stream into units (chunk runs starting at a dictionary reset) and decodes them in parallel.
type gpuReader2 struct {
	in       *bufio.Reader
	dictSize uint32
	out      []byte
	pos      int
	err      error
	decoded  bool
}

func (r *gpuReader2) decode() {
	data, _ := io.ReadAll(r.in)
	e, err := engine()
	if err != nil {
		r.err = err
		return
	}
	cunits, total, _ := ScanLZMA2(data, r.dictSize)
	units := make([]Unit, len(cunits))
	for i, u := range cunits {
		units[i] = Unit{Kind: 2, In: data[u.in_off : u.in_off+u.in_len], OutCap: uint64(u.out_cap),
			Lc: uint8(u.lc), Lp: uint8(u.lp), Pb: uint8(u.pb), DictSize: uint32(u.dict_size), Flags: uint32(u.flags)}
	}
	res, err := e.DecodeBatch(units)
	if err != nil {
		r.err = err
		return
	}
	r.out = make([]byte, 0, total)
	for _, x := range res { // deliver up to and including the first failing unit's prefix
		r.out = append(r.out, x.Out...)
		if x.Err != nil {
			r.err = x.Err
			break
		}
	}
}

func (r *gpuReader2) Read(p []byte) (int, error) {
	if !r.decoded {
		r.decode()
		r.decoded = true
	}
	n := copy(p, r.out[r.pos:])
	r.pos += n
	if n == len(p) && n > 0 {
		return n, nil
	}
	if r.err != nil {
		err := r.err
		r.err = nil
		return n, err
	}
	return n, io.EOF
}

// Drop-in replacements for the bodies of Reader1.Read / Reader2.Read (reader1.go:223-254,
// reader2.go:216-250): same exported API and error values, the in-stream decode runs on the GPU.
//
// NOT COMPILED HERE (no Go toolchain in the build image); see INTEGRATION.md. The same state machines
// exist in C++ (lzma_b200/csrc/lzma_reader.cpp) and Python (lzma_b200/reader1.py, reader2.py), which are
// compiled / run by the test-suite.
//
// A single .lzma stream is ONE unit = one warp of the GPU: it decodes at roughly 11 MB/s, several times
// slower than the CPU reader it replaces. The facade pays off for LZMA2 streams with dictionary resets
// (hundreds of units per wave) and for callers that batch (Engine.DecodeBatch / DecodeFolders).
package lzma

import (
	"io"
	"math"
	"sync"
)

var (
	defaultEngine    *Engine
	defaultEngineErr error
	defaultOnce      sync.Once
)

func engine() (*Engine, error) {
	defaultOnce.Do(func() { defaultEngine, defaultEngineErr = NewEngine() }) // no device: no fallback, the caller sees the error
	return defaultEngine, defaultEngineErr
}

// gpuReader1 keeps the constructor behaviour of NewReader1 (header + range-coder preamble are read
// and validated eagerly, reader1.go:77-159) and decodes the body on the first Read.
type gpuReader1 struct {
	in         io.ByteReader
	lc, lp, pb uint8
	dictSize   uint32
	unpackSize uint64
	preamble   [5]byte
	out        []byte
	pos        int
	err        error
	decoded    bool
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
	// The header's size field is untrusted (a 13-byte header can claim 2^50 bytes): the first capacity is
	// bounded by the payload, and grows towards the declared size only when the decoder asks for more.
	known := r.unpackSize != math.MaxUint64
	capacity := uint64(len(payload))*8 + 1<<16
	if known && r.unpackSize < capacity {
		capacity = r.unpackSize
	}
	for {
		res, err := e.DecodeBatch([]Unit{{Kind: KindLZMA1Raw, In: payload, OutCap: capacity, Lc: r.lc, Lp: r.lp, Pb: r.pb,
			DictSize: r.dictSize, UnpackSize: r.unpackSize}})
		if err != nil {
			r.err = err
			return
		}
		canGrow := capacity < 1<<40
		if known {
			canGrow = capacity < r.unpackSize
		}
		if res[0].Err == ErrOutputOverflow && canGrow {
			capacity *= 8 // the streaming API has no capacity: grow and decode again
			if known && capacity > r.unpackSize {
				capacity = r.unpackSize
			}
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

// gpuReader2: NewReader2's eager first-header read stays in the constructor. Read works in WAVES: chunk
// headers are walked on the host (reader2.go:100-214) until waveBytes of output are covered and the next
// chunk starts a unit that inherits nothing; that wave is decoded in one GPU call. While a wave is being
// served to the caller the next one is already being read and decoded by a goroutine (decode-ahead), so a
// steady reader sees the GPU's throughput, not decode + delivery in turns; memory is bounded by two waves.
type gpuReader2 struct {
	in        io.Reader
	dictSize  uint32
	buf       []byte // input read but not yet handed to the GPU
	inEOF     bool
	waveBytes uint64

	cur  *wave      // being served
	next chan *wave // decode-ahead result (capacity 1), nil until the first Read
}

type wave struct {
	out  []byte
	err  error
	last bool
}

const defaultWaveBytes = 1 << 30 // a wave takes as long as its longest unit (~100 ms per MiB of text): large waves give throughput

// fill makes at least need bytes available in r.buf[from:]; false if the input ended first.
func (r *gpuReader2) fill(from, need int) bool {
	for len(r.buf)-from < need && !r.inEOF {
		chunk := make([]byte, 1<<20)
		n, err := r.in.Read(chunk)
		r.buf = append(r.buf, chunk[:n]...)
		if err != nil || n == 0 {
			r.inEOF = true
		}
	}
	return len(r.buf)-from >= need
}

// independentFrom: an uncompressed dictionary-reset chunk at pos starts a unit that inherits nothing iff the
// first LZMA chunk after it (before the next reset) brings new properties (reader2.go:155-165).
func (r *gpuReader2) independentFrom(pos int) bool {
	p0 := pos
	for {
		if !r.fill(0, pos+3) {
			return true
		}
		c := r.buf[pos]
		if c == 0 || (c >= 3 && c < 0x80) || c >= 0xE0 || (c == 1 && pos != p0) {
			return true
		}
		if c >= 0x80 {
			return c >= 0xC0
		}
		pos += 3 + (int(r.buf[pos+1])<<8 | int(r.buf[pos+2])) + 1
	}
}

// cutWave returns the bytes of the next wave (terminated with 0x00 when it is not the stream's end) and
// whether it is the last one; the bytes are removed from r.buf.
func (r *gpuReader2) cutWave() ([]byte, bool) {
	pos, out, first := 0, uint64(0), true
	take := func(n int, term bool) []byte {
		w := append([]byte{}, r.buf[:n]...)
		if term {
			w = append(w, 0)
		}
		r.buf = r.buf[n:]
		return w
	}
	for {
		if !r.fill(0, pos+1) {
			return take(len(r.buf), false), true // ran off the input: the device reports it
		}
		ctrl := r.buf[pos]
		if ctrl == 0 || (ctrl >= 3 && ctrl < 0x80) { // end of stream (0x03-0x7F too, reader2.go:185-198)
			return take(pos+1, false), true
		}
		enough := !first && out >= r.waveBytes
		if enough && (ctrl >= 0xE0 || (ctrl == 1 && r.independentFrom(pos))) {
			return take(pos, true), false
		}
		hl := 3
		if ctrl >= 0xC0 {
			hl = 6
		} else if ctrl >= 0x80 {
			hl = 5
		}
		if !r.fill(0, pos+hl) {
			return take(len(r.buf), false), true
		}
		usz := (int(r.buf[pos+1])<<8 | int(r.buf[pos+2])) + 1
		payload := usz
		if ctrl >= 0x80 {
			usz += int(ctrl&0x1F) << 16
			payload = (int(r.buf[pos+3])<<8 | int(r.buf[pos+4])) + 1
		}
		if !r.fill(0, pos+hl+payload) {
			return take(len(r.buf), false), true
		}
		pos += hl + payload
		out += uint64(usz)
		first = false
	}
}

func (r *gpuReader2) decodeWave() *wave {
	data, last := r.cutWave()
	e, err := engine()
	if err != nil {
		return &wave{err: err, last: true}
	}
	units, total, _ := ScanLZMA2(data, r.dictSize)
	res, err := e.DecodeBatch(units)
	if err != nil {
		return &wave{err: err, last: true}
	}
	w := &wave{out: make([]byte, 0, total), last: last}
	for _, x := range res { // deliver up to and including the first failing unit's prefix
		w.out = append(w.out, x.Out...)
		if x.Err != nil {
			w.err, w.last = x.Err, true
			break
		}
	}
	return w
}

// ahead starts decoding the next wave in the background (only this goroutine touches r.in / r.buf until the
// result has been received).
func (r *gpuReader2) ahead() {
	r.next = make(chan *wave, 1)
	go func(ch chan *wave) { ch <- r.decodeWave() }(r.next)
}

func (r *gpuReader2) Read(p []byte) (int, error) {
	if r.waveBytes == 0 {
		r.waveBytes = defaultWaveBytes
	}
	if r.cur == nil {
		r.cur = r.decodeWave()
		if !r.cur.last {
			r.ahead()
		}
	}
	for len(r.cur.out) == 0 && !r.cur.last && len(p) > 0 { // previous wave delivered: take the one decoded meanwhile
		r.cur = <-r.next
		r.next = nil
		if !r.cur.last {
			r.ahead()
		}
	}
	n := copy(p, r.cur.out)
	r.cur.out = r.cur.out[n:]
	if (n == len(p) && n > 0) || !r.cur.last {
		return n, nil
	}
	if r.cur.err != nil {
		err := r.cur.err
		r.cur.err = nil
		return n, err
	}
	return n, io.EOF
}

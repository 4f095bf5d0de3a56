"""A small interpreter for the PTX subset of lzma_b200/csrc/lzgpu_fast2.cuh's bit ladders, and the extraction of those
asm blocks from the header (gcc -E over the macro section), so that the text that ships can run on a machine without a
GPU.  Test infrastructure: scalar (one lane), 32-bit registers, one flat shared-memory array."""
import os
import re
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
HEADER = os.path.join(ROOT, "lzma_b200", "csrc", "lzgpu_fast2.cuh")
M = 0xFFFFFFFF


# ----------------------------------------------------------------------------- extraction
def _macro_section():
    src = open(HEADER).read()
    a = src.index("// ---- PTX building blocks")
    b = src.index("#define F2_FAIL(ST, SITE)")
    return src[a:b]


def _split_top(s, sep):
    """split s at `sep` characters that are outside string literals and parentheses"""
    out, depth, cur, i, in_str = [], 0, [], 0, False
    while i < len(s):
        c = s[i]
        if in_str:
            cur.append(c)
            if c == "\\":
                cur.append(s[i + 1]); i += 1
            elif c == '"':
                in_str = False
        elif c == '"':
            in_str = True; cur.append(c)
        elif c in "([":
            depth += 1; cur.append(c)
        elif c in ")]":
            depth -= 1; cur.append(c)
        elif c == sep and depth == 0:
            out.append("".join(cur)); cur = []
        else:
            cur.append(c)
        i += 1
    out.append("".join(cur))
    return out


def _strings(s):
    parts = re.findall(r'"((?:[^"\\]|\\.)*)"', s)
    return "".join(p.encode().decode("unicode_escape") for p in parts)


class Block:
    def __init__(self, text, outs, ins):
        self.text, self.outs, self.ins = text, outs, ins      # outs / ins: [(constraint, expression)]
        self.prog, self.labels, self.regs = _compile(text)


def extract(invocations, defines=()):
    """invocations: {name: 'F2_BIT(D, PV, A, BIT)'} -> {name: Block}"""
    body = ["#define __CUDA_ARCH__ 1000"] + [f"#define {d}" for d in defines] + [_macro_section()]
    for name, inv in invocations.items():
        body.append(f"@@{name}@@ {inv} @@END@@")
    with tempfile.NamedTemporaryFile("w", suffix=".c", delete=False) as f:
        f.write("\n".join(body))
        path = f.name
    try:
        out = subprocess.run(["gcc", "-E", "-P", "-x", "c", path], capture_output=True, text=True, check=True).stdout
    finally:
        os.unlink(path)
    blocks = {}
    for name in invocations:
        m = re.search(r"@@%s@@(.*?)@@END@@" % re.escape(name), out, re.S)
        exp = m.group(1)
        inner = exp[exp.index("(") + 1:exp.rindex(")")]
        if inner.lstrip().startswith("volatile"):
            pass
        secs = _split_top(inner, ":")
        text = _strings(secs[0])

        def ops(sec):
            res = []
            for o in _split_top(sec, ","):
                o = o.strip()
                if not o:
                    continue
                mm = re.match(r'"([^"]*)"\s*\((.*)\)\s*$', o, re.S)
                res.append((mm.group(1), mm.group(2).strip()))
            return res
        outs = ops(secs[1]) if len(secs) > 1 else []
        ins = ops(secs[2]) if len(secs) > 2 else []
        blocks[name] = Block(text, outs, ins)
    return blocks


# ----------------------------------------------------------------------------- interpreter
def _compile(text):
    prog, labels, regs = [], {}, {}
    for chunk in text.split(";"):
        c = chunk.strip()
        while c:
            if c[0] in "{}":
                c = c[1:].strip(); continue
            m = re.match(r"^([A-Za-z_][A-Za-z_0-9]*):\s*", c)
            if m:
                labels[m.group(1)] = len(prog)
                c = c[m.end():].strip(); continue
            break
        if not c:
            continue
        if c.startswith(".reg"):
            kind, names = re.match(r"\.reg\s+(\.\w+)\s+(.*)$", c, re.S).groups()
            for n in names.split(","):
                n = n.strip()
                mm = re.match(r"^(\w+)<(\d+)>$", n)
                for r in ([f"{mm.group(1)}{i}" for i in range(int(mm.group(2)))] if mm else [n]):
                    regs[r] = False if kind == ".pred" else 0
            continue
        guard = None
        m = re.match(r"^@(!?)(\w+)\s+", c)
        if m:
            guard = (m.group(2), m.group(1) == "!")
            c = c[m.end():]
        op, _, rest = c.partition(" ")
        args = [a.strip() for a in _split_top(rest.replace("{", "").replace("}", ""), ",")] if rest.strip() else []
        prog.append((guard, op, args))
    return prog, labels, regs


class OutOfBounds(Exception):
    pass


class Shared:
    """Flat shared memory with bounds: [0, table_end) holds uint16 probability cells (the only place a block may store
    to, and where its 16-bit loads must lie), [in_lo, in_hi) the staged input (8-bit loads only)."""

    def __init__(self, size, table_end=None, in_lo=None, in_hi=None):
        self.b = bytearray(size)
        self.table_end, self.in_lo, self.in_hi = table_end, in_lo, in_hi

    def ld(self, a, n):
        if self.table_end is not None:
            ok = (a + n <= self.table_end and a % 2 == 0) if n == 2 else (self.in_lo <= a < self.in_hi)
            if not ok:
                raise OutOfBounds(f"{n}-byte load at {a:#x}")
        return int.from_bytes(self.b[a:a + n], "little")

    def st(self, a, n, v):
        if self.table_end is not None and not (n == 2 and a % 2 == 0 and a + 2 <= self.table_end):
            raise OutOfBounds(f"{n}-byte store at {a:#x}")
        self.b[a:a + n] = (v & ((1 << (8 * n)) - 1)).to_bytes(n, "little")


def _s32(v):
    v &= M
    return v - (1 << 32) if v & 0x80000000 else v


def run(block, values, sm, counters=None):
    """values: operand values in order outputs then inputs (python ints); returns the list after execution."""
    regs = dict(block.regs)
    vals = list(values)

    def get(a):
        if a.startswith("%"):
            return vals[int(a[1:])] & M
        if a in regs:
            return regs[a]
        return int(a, 0) & M

    def put(a, v):
        if a.startswith("%"):
            vals[int(a[1:])] = v & M
        else:
            assert a in regs, a
            regs[a] = v if isinstance(regs[a], bool) else v & M

    def addr(a):
        inner = a.strip()[1:-1]
        base, _, off = inner.partition("+")
        return (get(base.strip()) + (int(off, 0) if off else 0)) & M

    pc, steps, n = 0, 0, len(block.prog)
    while pc < n:
        guard, op, args = block.prog[pc]
        pc += 1
        steps += 1
        if guard is not None:
            p = regs[guard[0]]
            if p == guard[1]:
                continue
        if op.startswith("bra"):
            pc = block.labels[args[0]]
            continue
        base = op.split(".")[0]
        if base == "mov":
            put(args[0], get(args[1]))
        elif base == "shr":
            s = get(args[2])
            if op.endswith(".s32"):
                put(args[0], (_s32(get(args[1])) >> min(s, 31)) & M)
            else:
                put(args[0], (get(args[1]) >> s) if s < 32 else 0)
        elif base == "shl":
            s = get(args[2])
            put(args[0], (get(args[1]) << s) & M if s < 32 else 0)
        elif base == "add":
            put(args[0], get(args[1]) + get(args[2]))
        elif base == "sub":
            put(args[0], get(args[1]) - get(args[2]))
        elif base == "neg":
            put(args[0], -get(args[1]))
        elif base == "mul":
            assert ".lo" in op
            put(args[0], get(args[1]) * get(args[2]))
        elif base == "mad":
            assert ".lo" in op
            put(args[0], get(args[1]) * get(args[2]) + get(args[3]))
        elif base == "min":
            put(args[0], min(get(args[1]), get(args[2])))
        elif base == "and":
            put(args[0], get(args[1]) & get(args[2]))
        elif base == "or":
            put(args[0], get(args[1]) | get(args[2]))
        elif base == "xor":
            if op.endswith(".pred"):
                regs[args[0]] = regs[args[1]] != regs[args[2]]
            else:
                put(args[0], get(args[1]) ^ get(args[2]))
        elif base == "setp":
            cmp = op.split(".")[1]
            a, b = get(args[1]), get(args[2])
            r = {"ge": a >= b, "lt": a < b, "ne": a != b, "le": a <= b, "eq": a == b, "gt": a > b}[cmp]
            if ".and." in op:
                r = r and regs[args[3]]
            regs[args[0]] = bool(r)
        elif base == "selp":
            put(args[0], get(args[1]) if regs[args[3]] else get(args[2]))
        elif base == "ld":
            assert ".shared" in op
            put(args[0], sm.ld(addr(args[1]), 1 if op.endswith("u8") else 2))
        elif base == "st":
            assert ".shared" in op
            sm.st(addr(args[0]), 2 if op.endswith("u16") else 1, get(args[1]))
        else:
            raise NotImplementedError(op)
    if counters is not None:
        counters["steps"] = counters.get("steps", 0) + steps
    return vals

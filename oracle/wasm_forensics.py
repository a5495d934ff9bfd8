"""Minimal WebAssembly (MVP + bulk-memory/sat-trunc/sign-ext) parser and interpreter.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/gen_golden.py`` (run in the build
container, where ``/root/reference`` exists) to execute functions of the
reference's own shipped binary ``builds/web_build.zip:pkg/underwater_world_bg.wasm``
and record their outputs as golden vectors under ``tests/golden/``.  That binary
is the only executable artefact of the reference's arithmetic available without a
Rust toolchain.  Nothing in the product path imports
this module, and nothing at test time on the GPU box needs it (the vectors are
committed).

Written from the WebAssembly core specification; not derived from any code in
the reference.
"""
from __future__ import annotations

import math
import struct
import zipfile

MASK32 = 0xFFFFFFFF
MASK64 = 0xFFFFFFFFFFFFFFFF


class Trap(Exception):
    pass


class ImportCalled(Exception):
    def __init__(self, idx, name, args):
        super().__init__(f"import #{idx} {name} called")
        self.idx, self.name, self.args = idx, name, args


def _uleb(b, p):
    r = 0
    s = 0
    while True:
        x = b[p]
        p += 1
        r |= (x & 0x7F) << s
        s += 7
        if x < 0x80:
            return r, p


def _sleb(b, p):
    r = 0
    s = 0
    while True:
        x = b[p]
        p += 1
        r |= (x & 0x7F) << s
        s += 7
        if x < 0x80:
            if x & 0x40:
                r -= 1 << s
            return r, p


def _s32(x):
    x &= MASK32
    return x - (1 << 32) if x & 0x80000000 else x


def _s64(x):
    x &= MASK64
    return x - (1 << 64) if x & (1 << 63) else x


_pf = struct.Struct("<f")
_pd = struct.Struct("<d")
_pI = struct.Struct("<I")
_pQ = struct.Struct("<Q")


def f32r(x):
    """Round a Python float to the nearest binary32 (ties-to-even)."""
    try:
        return _pf.unpack(_pf.pack(x))[0]
    except OverflowError:
        return math.copysign(math.inf, x)


def _fmin(a, b):
    if a != a or b != b:
        return math.nan
    if a == 0.0 and b == 0.0:
        return a if math.copysign(1.0, a) < 0 else b
    return a if a < b else b


def _fmax(a, b):
    if a != a or b != b:
        return math.nan
    if a == 0.0 and b == 0.0:
        return a if math.copysign(1.0, a) > 0 else b
    return a if a > b else b


def _nearest(x):
    if x != x or math.isinf(x):
        return x
    r = round(x)  # Python rounds half to even
    return math.copysign(float(r), x) if r == 0 else float(r)


def _trunc_checked(x, lo, hi):
    if x != x:
        raise Trap("invalid conversion to integer")
    t = math.trunc(x)
    if t < lo or t > hi:
        raise Trap("integer overflow")
    return t


def _trunc_sat(x, lo, hi):
    if x != x:
        return 0
    if x == math.inf:
        return hi
    if x == -math.inf:
        return lo
    t = math.trunc(x)
    return lo if t < lo else hi if t > hi else t


def _clz(x, bits):
    return bits - x.bit_length()


def _ctz(x, bits):
    return bits if x == 0 else (x & -x).bit_length() - 1


class Func:
    __slots__ = ("type_idx", "locals", "code", "body_start", "body_size", "decoded")


class Module:
    """Parsed module.  Function index space = imports first, then defined."""

    def __init__(self, data: bytes):
        self.b = data
        assert data[:8] == b"\x00asm\x01\x00\x00\x00"
        self.types = []
        self.imports = []  # (module, name, kind, desc)
        self.func_imports = []  # indices into imports of kind func
        self.funcs = []  # defined
        self.tables = []
        self.mem_min = 0
        self.globals_init = []
        self.exports = {}
        self.elems = []
        self.datas = []
        self._parse()

    @classmethod
    def from_zip(cls, zip_path, member):
        with zipfile.ZipFile(zip_path) as z:
            return cls(z.read(member))

    def _const_expr(self, p):
        b = self.b
        op = b[p]
        p += 1
        if op == 0x41:
            v, p = _sleb(b, p)
            v &= MASK32
        elif op == 0x42:
            v, p = _sleb(b, p)
            v &= MASK64
        elif op == 0x43:
            v = _pf.unpack_from(b, p)[0]
            p += 4
        elif op == 0x44:
            v = _pd.unpack_from(b, p)[0]
            p += 8
        elif op == 0x23:
            gi, p = _uleb(b, p)
            v = ("global", gi)
        else:
            raise NotImplementedError(hex(op))
        assert b[p] == 0x0B
        return v, p + 1

    def _parse(self):
        b = self.b
        p = 8
        func_type_idx = []
        while p < len(b):
            sid = b[p]
            p += 1
            n, p = _uleb(b, p)
            end = p + n
            if sid == 1:
                cnt, q = _uleb(b, p)
                for _ in range(cnt):
                    assert b[q] == 0x60
                    q += 1
                    np_, q = _uleb(b, q)
                    params = list(b[q:q + np_])
                    q += np_
                    nr, q = _uleb(b, q)
                    results = list(b[q:q + nr])
                    q += nr
                    self.types.append((params, results))
            elif sid == 2:
                cnt, q = _uleb(b, p)
                for _ in range(cnt):
                    ln, q = _uleb(b, q)
                    mod = b[q:q + ln].decode()
                    q += ln
                    ln, q = _uleb(b, q)
                    nm = b[q:q + ln].decode()
                    q += ln
                    kind = b[q]
                    q += 1
                    if kind == 0:
                        ti, q = _uleb(b, q)
                        self.func_imports.append(len(self.imports))
                        self.imports.append((mod, nm, kind, ti))
                    elif kind == 1:
                        q += 1
                        fl, q = _uleb(b, q)
                        _, q = _uleb(b, q)
                        if fl & 1:
                            _, q = _uleb(b, q)
                        self.imports.append((mod, nm, kind, None))
                    elif kind == 2:
                        fl, q = _uleb(b, q)
                        _, q = _uleb(b, q)
                        if fl & 1:
                            _, q = _uleb(b, q)
                        self.imports.append((mod, nm, kind, None))
                    elif kind == 3:
                        q += 2
                        self.imports.append((mod, nm, kind, None))
            elif sid == 3:
                cnt, q = _uleb(b, p)
                for _ in range(cnt):
                    ti, q = _uleb(b, q)
                    func_type_idx.append(ti)
            elif sid == 4:
                cnt, q = _uleb(b, p)
                for _ in range(cnt):
                    q += 1
                    fl, q = _uleb(b, q)
                    mn, q = _uleb(b, q)
                    if fl & 1:
                        _, q = _uleb(b, q)
                    self.tables.append(mn)
            elif sid == 5:
                cnt, q = _uleb(b, p)
                fl, q = _uleb(b, q)
                self.mem_min, q = _uleb(b, q)
            elif sid == 6:
                cnt, q = _uleb(b, p)
                for _ in range(cnt):
                    q += 2
                    v, q = self._const_expr(q)
                    self.globals_init.append(v)
            elif sid == 7:
                cnt, q = _uleb(b, p)
                for _ in range(cnt):
                    ln, q = _uleb(b, q)
                    nm = b[q:q + ln].decode()
                    q += ln
                    kind = b[q]
                    q += 1
                    idx, q = _uleb(b, q)
                    self.exports[nm] = (kind, idx)
            elif sid == 9:
                cnt, q = _uleb(b, p)
                for _ in range(cnt):
                    fl, q = _uleb(b, q)
                    if fl != 0:
                        raise NotImplementedError("elem flags %d" % fl)
                    off, q = self._const_expr(q)
                    m, q = _uleb(b, q)
                    fs = []
                    for _ in range(m):
                        fi, q = _uleb(b, q)
                        fs.append(fi)
                    self.elems.append((off, fs))
            elif sid == 10:
                cnt, q = _uleb(b, p)
                for i in range(cnt):
                    sz, q = _uleb(b, q)
                    f = Func()
                    f.type_idx = func_type_idx[i]
                    f.body_start = q
                    f.body_size = sz
                    f.decoded = None
                    self.funcs.append(f)
                    q += sz
            elif sid == 11:
                cnt, q = _uleb(b, p)
                for _ in range(cnt):
                    fl, q = _uleb(b, q)
                    if fl == 0:
                        off, q = self._const_expr(q)
                    elif fl == 1:
                        off = None
                    else:
                        _, q = _uleb(b, q)
                        off, q = self._const_expr(q)
                    ln, q = _uleb(b, q)
                    self.datas.append((off, q, ln))
                    q += ln
            p = end
        self.n_func_imports = len(self.func_imports)

    # ------------------------------------------------------------------
    def decode(self, fidx):
        """Decode defined-or-imported function index `fidx` (function index space)."""
        f = self.funcs[fidx - self.n_func_imports]
        if f.decoded is not None:
            return f
        b = self.b
        p = f.body_start
        end = p + f.body_size
        nl, p = _uleb(b, p)
        locs = []
        for _ in range(nl):
            c, p = _uleb(b, p)
            t = b[p]
            p += 1
            locs.extend([t] * c)
        f.locals = locs
        code = []
        offs = []
        ctl = []  # stack of indices of block/loop/if
        while p < end:
            at = p
            op = b[p]
            p += 1
            imm = None
            if op in (0x02, 0x03, 0x04):
                bt, p = _sleb(b, p)
                imm = [bt, None, None]  # blocktype, end_idx, else_idx
                ctl.append(len(code))
            elif op == 0x05:
                code[ctl[-1]][1][2] = len(code)
                imm = ctl[-1]
            elif op == 0x0B:
                if ctl:
                    o = ctl.pop()
                    code[o][1][1] = len(code)
            elif op in (0x0C, 0x0D, 0x10, 0x20, 0x21, 0x22, 0x23, 0x24):
                imm, p = _uleb(b, p)
            elif op == 0x0E:
                n, p = _uleb(b, p)
                tg = []
                for _ in range(n + 1):
                    t, p = _uleb(b, p)
                    tg.append(t)
                imm = tg
            elif op == 0x11:
                ti, p = _uleb(b, p)
                tb, p = _uleb(b, p)
                imm = ti
            elif 0x28 <= op <= 0x3E:
                _, p = _uleb(b, p)
                imm, p = _uleb(b, p)
            elif op in (0x3F, 0x40):
                p += 1
            elif op == 0x41:
                v, p = _sleb(b, p)
                imm = v & MASK32
            elif op == 0x42:
                v, p = _sleb(b, p)
                imm = v & MASK64
            elif op == 0x43:
                imm = _pf.unpack_from(b, p)[0]
                p += 4
            elif op == 0x44:
                imm = _pd.unpack_from(b, p)[0]
                p += 8
            elif op == 0xFC:
                sub, p = _uleb(b, p)
                op = 0xFC00 | sub
                if sub == 10:
                    p += 2
                elif sub == 11:
                    p += 1
                elif sub in (8, 12, 14):
                    raise NotImplementedError("0xFC %d" % sub)
            code.append((op, imm))
            offs.append(at)
        f.code = code
        f.decoded = offs
        return f


class Instance:
    def __init__(self, mod: Module, extra_pages=64):
        self.m = mod
        self.mem = bytearray((mod.mem_min + extra_pages) * 65536)
        self.pages = mod.mem_min + extra_pages
        self.globals = []
        for g in mod.globals_init:
            self.globals.append(g)
        self.table = [None] * (mod.tables[0] if mod.tables else 0)
        for off, fs in mod.elems:
            for i, fi in enumerate(fs):
                self.table[off + i] = fi
        for off, q, ln in mod.datas:
            if off is not None:
                self.mem[off:off + ln] = mod.b[q:q + ln]
        self.import_hook = None  # callable(idx, name, args) -> list of results
        self.call_hook = None    # callable(fidx, args) -> None or list (to override)
        self.steps = 0

    # memory helpers ----------------------------------------------------
    def read(self, addr, n):
        return bytes(self.mem[addr:addr + n])

    def write(self, addr, data):
        self.mem[addr:addr + len(data)] = data

    def invoke(self, fidx, args):
        m = self.m
        if fidx < m.n_func_imports:
            mod, nm, _, ti = m.imports[m.func_imports[fidx]]
            if self.import_hook is None:
                raise ImportCalled(fidx, f"{mod}.{nm}", list(args))
            return self.import_hook(fidx, f"{mod}.{nm}", list(args))
        if self.call_hook is not None:
            r = self.call_hook(fidx, list(args))
            if r is not None:
                return r
        f = m.decode(fidx)
        params, results = m.types[f.type_idx]
        assert len(args) == len(params), (fidx, len(args), len(params))
        loc = list(args)
        for t in f.locals:
            loc.append(0.0 if t in (0x7D, 0x7C) else 0)
        return self._run(f, loc, len(results))

    def _arity(self, bt, is_loop):
        if bt == -64:  # 0x40 empty
            return 0
        if bt < 0:
            return 0 if is_loop else 1
        params, results = self.m.types[bt]
        return len(params) if is_loop else len(results)

    def run_region(self, fidx, start_idx, stop_idx, locals_init):
        """Execute instructions [start_idx, stop_idx) of a function body with preset locals.

        Used to run code that the compiler inlined into a larger function (e.g. the
        permutation-table construction inside State::new).  The region must open every
        block it branches out of.  Returns the locals list after execution.
        """
        f = self.m.decode(fidx)
        params, _ = self.m.types[f.type_idx]
        loc = [0] * len(params)
        for t in f.locals:
            loc.append(0.0 if t in (0x7D, 0x7C) else 0)
        for k, v in locals_init.items():
            loc[k] = v
        self._run(f, loc, 0, start_idx, stop_idx)
        return loc

    def _run(self, f, loc, nres, pc=0, stop=-1):
        code = f.code
        mem = self.mem
        st = []
        push = st.append
        pop = st.pop
        # control stack entries: (branch_target_pc, stack_height, arity)
        cs = [(len(code), 0, nres)]
        n = len(code)
        steps = 0
        while pc < n:
            if pc == stop:
                break
            op, imm = code[pc]
            pc += 1
            steps += 1
            if op == 0x20:
                push(loc[imm])
            elif op == 0x21:
                loc[imm] = pop()
            elif op == 0x22:
                loc[imm] = st[-1]
            elif op == 0x41 or op == 0x42 or op == 0x43 or op == 0x44:
                push(imm)
            elif op == 0xA0:
                b_ = pop(); st[-1] = st[-1] + b_
            elif op == 0xA1:
                b_ = pop(); st[-1] = st[-1] - b_
            elif op == 0xA2:
                b_ = pop(); st[-1] = st[-1] * b_
            elif op == 0x6A:
                b_ = pop(); st[-1] = (st[-1] + b_) & MASK32
            elif op == 0x6B:
                b_ = pop(); st[-1] = (st[-1] - b_) & MASK32
            elif op == 0x71:
                b_ = pop(); st[-1] = st[-1] & b_
            elif op == 0x72:
                b_ = pop(); st[-1] = st[-1] | b_
            elif op == 0x73:
                b_ = pop(); st[-1] = st[-1] ^ b_
            elif op == 0x02:
                cs.append((imm[1] + 1, len(st), self._arity(imm[0], False)))
            elif op == 0x03:
                cs.append((pc, len(st), self._arity(imm[0], True), True))
            elif op == 0x04:
                c = pop()
                cs.append((imm[1] + 1, len(st), self._arity(imm[0], False)))
                if not c:
                    if imm[2] is not None:
                        pc = imm[2] + 1
                    else:
                        pc = imm[1] + 1
                        cs.pop()
            elif op == 0x05:
                # reached else from then-branch: jump to end
                blk = code[imm][1]
                pc = blk[1] + 1
                cs.pop()
            elif op == 0x0B:
                cs.pop()
            elif op == 0x0C or op == 0x0D or op == 0x0E:
                if op == 0x0D:
                    if not pop():
                        continue
                    d = imm
                elif op == 0x0E:
                    i = pop()
                    d = imm[i] if i < len(imm) - 1 else imm[-1]
                else:
                    d = imm
                tgt = cs[-1 - d]
                ar = tgt[2]
                if ar:
                    vals = st[-ar:]
                    del st[tgt[1]:]
                    st.extend(vals)
                else:
                    del st[tgt[1]:]
                pc = tgt[0]
                if len(tgt) == 4:  # loop: keep its frame
                    del cs[len(cs) - d:]
                else:
                    del cs[len(cs) - 1 - d:]
            elif op == 0x0F:
                break
            elif op == 0x10:
                fi = imm
                m = self.m
                if fi < m.n_func_imports:
                    ti = m.imports[m.func_imports[fi]][3]
                else:
                    ti = m.funcs[fi - m.n_func_imports].type_idx
                npar = len(m.types[ti][0])
                if npar:
                    args = st[-npar:]
                    del st[-npar:]
                else:
                    args = []
                self.steps += steps
                steps = 0
                st.extend(self.invoke(fi, args))
            elif op == 0x11:
                ti = imm
                idx = pop()
                fi = self.table[idx] if idx < len(self.table) else None
                if fi is None:
                    raise Trap("undefined table element %d" % idx)
                npar = len(self.m.types[ti][0])
                if npar:
                    args = st[-npar:]
                    del st[-npar:]
                else:
                    args = []
                self.steps += steps
                steps = 0
                st.extend(self.invoke(fi, args))
            elif op == 0x00:
                raise Trap("unreachable executed")
            elif op == 0x01:
                pass
            elif op == 0x1A:
                pop()
            elif op == 0x1B:
                c = pop(); b_ = pop()
                if not c:
                    st[-1] = b_
            elif op == 0x23:
                push(self.globals[imm])
            elif op == 0x24:
                self.globals[imm] = pop()
            # ---- loads
            elif op == 0x28:
                a = st[-1] + imm; st[-1] = _pI.unpack_from(mem, a)[0]
            elif op == 0x29:
                a = st[-1] + imm; st[-1] = _pQ.unpack_from(mem, a)[0]
            elif op == 0x2A:
                a = st[-1] + imm; st[-1] = _pf.unpack_from(mem, a)[0]
            elif op == 0x2B:
                a = st[-1] + imm; st[-1] = _pd.unpack_from(mem, a)[0]
            elif op == 0x2C:
                a = st[-1] + imm; v = mem[a]; st[-1] = (v - 256 if v & 0x80 else v) & MASK32
            elif op == 0x2D:
                a = st[-1] + imm; st[-1] = mem[a]
            elif op == 0x2E:
                a = st[-1] + imm; v = mem[a] | (mem[a + 1] << 8); st[-1] = (v - 65536 if v & 0x8000 else v) & MASK32
            elif op == 0x2F:
                a = st[-1] + imm; st[-1] = mem[a] | (mem[a + 1] << 8)
            elif op == 0x30:
                a = st[-1] + imm; v = mem[a]; st[-1] = (v - 256 if v & 0x80 else v) & MASK64
            elif op == 0x31:
                a = st[-1] + imm; st[-1] = mem[a]
            elif op == 0x32:
                a = st[-1] + imm; v = mem[a] | (mem[a + 1] << 8); st[-1] = (v - 65536 if v & 0x8000 else v) & MASK64
            elif op == 0x33:
                a = st[-1] + imm; st[-1] = mem[a] | (mem[a + 1] << 8)
            elif op == 0x34:
                a = st[-1] + imm; st[-1] = _s32(_pI.unpack_from(mem, a)[0]) & MASK64
            elif op == 0x35:
                a = st[-1] + imm; st[-1] = _pI.unpack_from(mem, a)[0]
            # ---- stores
            elif op == 0x36:
                v = pop(); a = pop() + imm; _pI.pack_into(mem, a, v & MASK32)
            elif op == 0x37:
                v = pop(); a = pop() + imm; _pQ.pack_into(mem, a, v & MASK64)
            elif op == 0x38:
                v = pop(); a = pop() + imm; _pf.pack_into(mem, a, v)
            elif op == 0x39:
                v = pop(); a = pop() + imm; _pd.pack_into(mem, a, v)
            elif op == 0x3A or op == 0x3C:
                v = pop(); a = pop() + imm; mem[a] = v & 0xFF
            elif op == 0x3B or op == 0x3D:
                v = pop(); a = pop() + imm; mem[a] = v & 0xFF; mem[a + 1] = (v >> 8) & 0xFF
            elif op == 0x3E:
                v = pop(); a = pop() + imm; _pI.pack_into(mem, a, v & MASK32)
            elif op == 0x3F:
                push(self.pages)
            elif op == 0x40:
                d = pop()
                old = self.pages
                if old + d > 16384:
                    push(MASK32)
                else:
                    mem.extend(bytes(d * 65536))
                    self.pages += d
                    push(old)
            # ---- i32 compare
            elif op == 0x45:
                st[-1] = 1 if st[-1] == 0 else 0
            elif op == 0x46:
                b_ = pop(); st[-1] = 1 if st[-1] == b_ else 0
            elif op == 0x47:
                b_ = pop(); st[-1] = 1 if st[-1] != b_ else 0
            elif op == 0x48:
                b_ = pop(); st[-1] = 1 if _s32(st[-1]) < _s32(b_) else 0
            elif op == 0x49:
                b_ = pop(); st[-1] = 1 if st[-1] < b_ else 0
            elif op == 0x4A:
                b_ = pop(); st[-1] = 1 if _s32(st[-1]) > _s32(b_) else 0
            elif op == 0x4B:
                b_ = pop(); st[-1] = 1 if st[-1] > b_ else 0
            elif op == 0x4C:
                b_ = pop(); st[-1] = 1 if _s32(st[-1]) <= _s32(b_) else 0
            elif op == 0x4D:
                b_ = pop(); st[-1] = 1 if st[-1] <= b_ else 0
            elif op == 0x4E:
                b_ = pop(); st[-1] = 1 if _s32(st[-1]) >= _s32(b_) else 0
            elif op == 0x4F:
                b_ = pop(); st[-1] = 1 if st[-1] >= b_ else 0
            # ---- i64 compare
            elif op == 0x50:
                st[-1] = 1 if st[-1] == 0 else 0
            elif op == 0x51:
                b_ = pop(); st[-1] = 1 if st[-1] == b_ else 0
            elif op == 0x52:
                b_ = pop(); st[-1] = 1 if st[-1] != b_ else 0
            elif op == 0x53:
                b_ = pop(); st[-1] = 1 if _s64(st[-1]) < _s64(b_) else 0
            elif op == 0x54:
                b_ = pop(); st[-1] = 1 if st[-1] < b_ else 0
            elif op == 0x55:
                b_ = pop(); st[-1] = 1 if _s64(st[-1]) > _s64(b_) else 0
            elif op == 0x56:
                b_ = pop(); st[-1] = 1 if st[-1] > b_ else 0
            elif op == 0x57:
                b_ = pop(); st[-1] = 1 if _s64(st[-1]) <= _s64(b_) else 0
            elif op == 0x58:
                b_ = pop(); st[-1] = 1 if st[-1] <= b_ else 0
            elif op == 0x59:
                b_ = pop(); st[-1] = 1 if _s64(st[-1]) >= _s64(b_) else 0
            elif op == 0x5A:
                b_ = pop(); st[-1] = 1 if st[-1] >= b_ else 0
            # ---- float compare (f32 and f64 share Python semantics)
            elif op == 0x5B or op == 0x61:
                b_ = pop(); st[-1] = 1 if st[-1] == b_ else 0
            elif op == 0x5C or op == 0x62:
                b_ = pop(); st[-1] = 1 if st[-1] != b_ else 0
            elif op == 0x5D or op == 0x63:
                b_ = pop(); st[-1] = 1 if st[-1] < b_ else 0
            elif op == 0x5E or op == 0x64:
                b_ = pop(); st[-1] = 1 if st[-1] > b_ else 0
            elif op == 0x5F or op == 0x65:
                b_ = pop(); st[-1] = 1 if st[-1] <= b_ else 0
            elif op == 0x60 or op == 0x66:
                b_ = pop(); st[-1] = 1 if st[-1] >= b_ else 0
            # ---- i32 arith
            elif op == 0x67:
                st[-1] = _clz(st[-1], 32)
            elif op == 0x68:
                st[-1] = _ctz(st[-1], 32)
            elif op == 0x69:
                st[-1] = bin(st[-1]).count("1")
            elif op == 0x6C:
                b_ = pop(); st[-1] = (st[-1] * b_) & MASK32
            elif op == 0x6D:
                b_ = _s32(pop()); a = _s32(st[-1])
                if b_ == 0:
                    raise Trap("integer divide by zero")
                q = abs(a) // abs(b_)
                if (a < 0) != (b_ < 0):
                    q = -q
                if q > 0x7FFFFFFF:
                    raise Trap("integer overflow")
                st[-1] = q & MASK32
            elif op == 0x6E:
                b_ = pop()
                if b_ == 0:
                    raise Trap("integer divide by zero")
                st[-1] = st[-1] // b_
            elif op == 0x6F:
                b_ = _s32(pop()); a = _s32(st[-1])
                if b_ == 0:
                    raise Trap("integer divide by zero")
                r = abs(a) % abs(b_)
                st[-1] = (-r if a < 0 else r) & MASK32
            elif op == 0x70:
                b_ = pop()
                if b_ == 0:
                    raise Trap("integer divide by zero")
                st[-1] = st[-1] % b_
            elif op == 0x74:
                b_ = pop(); st[-1] = (st[-1] << (b_ & 31)) & MASK32
            elif op == 0x75:
                b_ = pop(); st[-1] = (_s32(st[-1]) >> (b_ & 31)) & MASK32
            elif op == 0x76:
                b_ = pop(); st[-1] = st[-1] >> (b_ & 31)
            elif op == 0x77:
                b_ = pop() & 31; a = st[-1]; st[-1] = ((a << b_) | (a >> (32 - b_))) & MASK32
            elif op == 0x78:
                b_ = pop() & 31; a = st[-1]; st[-1] = ((a >> b_) | (a << (32 - b_))) & MASK32
            # ---- i64 arith
            elif op == 0x79:
                st[-1] = _clz(st[-1], 64)
            elif op == 0x7A:
                st[-1] = _ctz(st[-1], 64)
            elif op == 0x7B:
                st[-1] = bin(st[-1]).count("1")
            elif op == 0x7C:
                b_ = pop(); st[-1] = (st[-1] + b_) & MASK64
            elif op == 0x7D:
                b_ = pop(); st[-1] = (st[-1] - b_) & MASK64
            elif op == 0x7E:
                b_ = pop(); st[-1] = (st[-1] * b_) & MASK64
            elif op == 0x7F:
                b_ = _s64(pop()); a = _s64(st[-1])
                if b_ == 0:
                    raise Trap("integer divide by zero")
                q = abs(a) // abs(b_)
                if (a < 0) != (b_ < 0):
                    q = -q
                st[-1] = q & MASK64
            elif op == 0x80:
                b_ = pop()
                if b_ == 0:
                    raise Trap("integer divide by zero")
                st[-1] = st[-1] // b_
            elif op == 0x81:
                b_ = _s64(pop()); a = _s64(st[-1])
                if b_ == 0:
                    raise Trap("integer divide by zero")
                r = abs(a) % abs(b_)
                st[-1] = (-r if a < 0 else r) & MASK64
            elif op == 0x82:
                b_ = pop()
                if b_ == 0:
                    raise Trap("integer divide by zero")
                st[-1] = st[-1] % b_
            elif op == 0x83:
                b_ = pop(); st[-1] = st[-1] & b_
            elif op == 0x84:
                b_ = pop(); st[-1] = st[-1] | b_
            elif op == 0x85:
                b_ = pop(); st[-1] = st[-1] ^ b_
            elif op == 0x86:
                b_ = pop(); st[-1] = (st[-1] << (b_ & 63)) & MASK64
            elif op == 0x87:
                b_ = pop(); st[-1] = (_s64(st[-1]) >> (b_ & 63)) & MASK64
            elif op == 0x88:
                b_ = pop(); st[-1] = st[-1] >> (b_ & 63)
            elif op == 0x89:
                b_ = pop() & 63; a = st[-1]; st[-1] = ((a << b_) | (a >> (64 - b_))) & MASK64
            elif op == 0x8A:
                b_ = pop() & 63; a = st[-1]; st[-1] = ((a >> b_) | (a << (64 - b_))) & MASK64
            # ---- f32 arith (round after every op)
            elif op == 0x8B:
                st[-1] = abs(st[-1])
            elif op == 0x8C:
                st[-1] = -st[-1]
            elif op == 0x8D:
                st[-1] = float(math.ceil(st[-1])) if math.isfinite(st[-1]) else st[-1]
            elif op == 0x8E:
                st[-1] = float(math.floor(st[-1])) if math.isfinite(st[-1]) else st[-1]
            elif op == 0x8F:
                st[-1] = float(math.trunc(st[-1])) if math.isfinite(st[-1]) else st[-1]
            elif op == 0x90:
                st[-1] = _nearest(st[-1])
            elif op == 0x91:
                st[-1] = f32r(math.sqrt(st[-1])) if st[-1] >= 0 else math.nan
            elif op == 0x92:
                b_ = pop(); st[-1] = f32r(st[-1] + b_)
            elif op == 0x93:
                b_ = pop(); st[-1] = f32r(st[-1] - b_)
            elif op == 0x94:
                b_ = pop(); st[-1] = f32r(st[-1] * b_)
            elif op == 0x95:
                b_ = pop(); a = st[-1]
                # double division of two binary32 values then rounding to binary32
                # is correctly rounded (2p+2 <= 53).
                if b_ == 0.0:
                    st[-1] = math.nan if (a == 0.0 or a != a) else math.copysign(math.inf, a) * math.copysign(1.0, b_)
                else:
                    st[-1] = f32r(a / b_)
            elif op == 0x96:
                b_ = pop(); st[-1] = _fmin(st[-1], b_)
            elif op == 0x97:
                b_ = pop(); st[-1] = _fmax(st[-1], b_)
            elif op == 0x98 or op == 0xA6:
                b_ = pop(); st[-1] = math.copysign(st[-1], b_)
            # ---- f64 arith
            elif op == 0x99:
                st[-1] = abs(st[-1])
            elif op == 0x9A:
                st[-1] = -st[-1]
            elif op == 0x9B:
                st[-1] = float(math.ceil(st[-1])) if math.isfinite(st[-1]) else st[-1]
            elif op == 0x9C:
                v = st[-1]
                if math.isfinite(v):
                    r = float(math.floor(v))
                    st[-1] = math.copysign(r, v) if r == 0 else r
            elif op == 0x9D:
                v = st[-1]
                if math.isfinite(v):
                    r = float(math.trunc(v))
                    st[-1] = math.copysign(r, v) if r == 0 else r
            elif op == 0x9E:
                st[-1] = _nearest(st[-1])
            elif op == 0x9F:
                st[-1] = math.sqrt(st[-1]) if st[-1] >= 0 else math.nan
            elif op == 0xA3:
                b_ = pop(); a = st[-1]
                if b_ == 0.0:
                    st[-1] = math.nan if (a == 0.0 or a != a) else math.copysign(math.inf, a) * math.copysign(1.0, b_)
                else:
                    st[-1] = a / b_
            elif op == 0xA4:
                b_ = pop(); st[-1] = _fmin(st[-1], b_)
            elif op == 0xA5:
                b_ = pop(); st[-1] = _fmax(st[-1], b_)
            # ---- conversions
            elif op == 0xA7:
                st[-1] = st[-1] & MASK32
            elif op == 0xA8 or op == 0xAA:
                st[-1] = _trunc_checked(st[-1], -(1 << 31), (1 << 31) - 1) & MASK32
            elif op == 0xA9 or op == 0xAB:
                st[-1] = _trunc_checked(st[-1], 0, MASK32)
            elif op == 0xAC:
                st[-1] = _s32(st[-1]) & MASK64
            elif op == 0xAD:
                pass
            elif op == 0xAE or op == 0xB0:
                st[-1] = _trunc_checked(st[-1], -(1 << 63), (1 << 63) - 1) & MASK64
            elif op == 0xAF or op == 0xB1:
                st[-1] = _trunc_checked(st[-1], 0, MASK64)
            elif op == 0xB2:
                st[-1] = f32r(float(_s32(st[-1])))
            elif op == 0xB3:
                st[-1] = f32r(float(st[-1]))
            elif op == 0xB4:
                st[-1] = _int_to_f32(_s64(st[-1]))
            elif op == 0xB5:
                st[-1] = _int_to_f32(st[-1])
            elif op == 0xB6:
                st[-1] = f32r(st[-1])
            elif op == 0xB7:
                st[-1] = float(_s32(st[-1]))
            elif op == 0xB8:
                st[-1] = float(st[-1])
            elif op == 0xB9:
                st[-1] = float(_s64(st[-1]))
            elif op == 0xBA:
                st[-1] = float(st[-1])
            elif op == 0xBB:
                pass
            elif op == 0xBC:
                st[-1] = _pI.unpack(_pf.pack(st[-1]))[0]
            elif op == 0xBD:
                st[-1] = _pQ.unpack(_pd.pack(st[-1]))[0]
            elif op == 0xBE:
                st[-1] = _pf.unpack(_pI.pack(st[-1]))[0]
            elif op == 0xBF:
                st[-1] = _pd.unpack(_pQ.pack(st[-1]))[0]
            elif op == 0xC0:
                v = st[-1] & 0xFF; st[-1] = (v - 256 if v & 0x80 else v) & MASK32
            elif op == 0xC1:
                v = st[-1] & 0xFFFF; st[-1] = (v - 65536 if v & 0x8000 else v) & MASK32
            elif op == 0xC2:
                v = st[-1] & 0xFF; st[-1] = (v - 256 if v & 0x80 else v) & MASK64
            elif op == 0xC3:
                v = st[-1] & 0xFFFF; st[-1] = (v - 65536 if v & 0x8000 else v) & MASK64
            elif op == 0xC4:
                st[-1] = _s32(st[-1]) & MASK64
            elif op >= 0xFC00:
                sub = op & 0xFF
                if sub in (0, 2):
                    st[-1] = _trunc_sat(st[-1], -(1 << 31), (1 << 31) - 1) & MASK32
                elif sub in (1, 3):
                    st[-1] = _trunc_sat(st[-1], 0, MASK32)
                elif sub in (4, 6):
                    st[-1] = _trunc_sat(st[-1], -(1 << 63), (1 << 63) - 1) & MASK64
                elif sub in (5, 7):
                    st[-1] = _trunc_sat(st[-1], 0, MASK64)
                elif sub == 10:
                    nbytes = pop(); src = pop(); dst = pop()
                    mem[dst:dst + nbytes] = mem[src:src + nbytes]
                elif sub == 11:
                    nbytes = pop(); val = pop(); dst = pop()
                    mem[dst:dst + nbytes] = bytes([val & 0xFF]) * nbytes
                else:
                    raise NotImplementedError(hex(op))
            else:
                raise NotImplementedError("opcode 0x%x" % op)
        self.steps += steps
        return st[len(st) - nres:] if nres else []


def _int_to_f32(i):
    # exact integer -> binary32, single rounding (avoid double rounding via float())
    if i == 0:
        return 0.0
    neg = i < 0
    a = -i if neg else i
    nb = a.bit_length()
    if nb <= 24:
        r = float(a)
    else:
        sh = nb - 24
        q = a >> sh
        rem = a & ((1 << sh) - 1)
        half = 1 << (sh - 1)
        if rem > half or (rem == half and (q & 1)):
            q += 1
        r = float(q) * (2.0 ** sh)
    return -r if neg else r


REF_ZIP = "/root/reference/builds/web_build.zip"
REF_WASM = "pkg/underwater_world_bg.wasm"


def load_reference_module():
    return Module.from_zip(REF_ZIP, REF_WASM)

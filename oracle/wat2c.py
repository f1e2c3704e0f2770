#!/usr/bin/env python3
"""WAT -> C transpiler for the reference's five FFT modules (TEST INFRASTRUCTURE).

This is oracle tooling, not product code.  It reads the reference's
``modules/*.wat`` *where they lie* (``/root/reference/modules``) and writes
generated C **only** into ``oracle/_ref/`` (git-ignored), from which
``oracle/Makefile`` builds ``oracle/_ref/libwatref.so``: the reference's own
kernels compiled natively.  No reference source is copied into the repo.

The modules use a small folded-S-expression subset of WebAssembly text
(SURVEY.md Appendix D): no imports, tables, data segments, memargs, typed
blocks, ``local.tee`` or ``select``.  Semantics preserved here:

* i32 ops are ``uint32_t`` arithmetic (wrap-around); shifts mask the count.
* f32/f64 scalar and lane ops are IEEE-754 round-to-nearest; the C must be
  built with ``-ffp-contract=off`` because WebAssembly has no fused
  multiply-add.  Lane ops use GCC vector extensions (per-lane IEEE).
* float constants are emitted as decimal text with the target-type suffix so
  the C compiler rounds decimal -> f32 directly (no double rounding).
* linear memory is a caller-owned byte block passed as the first argument of
  every function, so one "instance" == one memory block (thread-safe, like one
  WebAssembly.Instance per worker).

Generated symbols for module ``foo`` with export ``bar(i32)``:
``void watref_foo_bar(uint8_t *mem, uint32_t n)``;
``uint32_t watref_foo_pages(void)``; exported i32 globals become
``uint32_t watref_foo_global_NAME(void)``.
"""
from __future__ import annotations

import re
import sys
from pathlib import Path

# --------------------------------------------------------------------------
# S-expression reader
# --------------------------------------------------------------------------
_TOKEN = re.compile(r'"(?:[^"\\]|\\.)*"|[()]|[^\s()]+')


def read_sexprs(text: str):
    text = re.sub(r";;[^\n]*", "", text)
    toks = _TOKEN.findall(text)
    pos = 0

    def parse():
        nonlocal pos
        t = toks[pos]
        pos += 1
        if t == "(":
            lst = []
            while toks[pos] != ")":
                lst.append(parse())
            pos += 1
            return lst
        if t == ")":
            raise SyntaxError("unbalanced )")
        return t

    out = []
    while pos < len(toks):
        out.append(parse())
    return out


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
CTYPE = {"i32": "uint32_t", "f32": "float", "f64": "double", "v128": "v128"}


def cident(name: str) -> str:
    return re.sub(r"[^A-Za-z0-9_]", "_", name.lstrip("$"))


def fconst(txt: str, ty: str) -> str:
    t = txt.replace("_", "")
    if t in ("inf", "+inf"):
        return "INFINITY"
    if t == "-inf":
        return "(-INFINITY)"
    if "nan" in t:
        raise NotImplementedError("nan constants")
    if t.lower().startswith(("0x", "-0x", "+0x")):
        if "p" not in t.lower():
            t += "p0"
    elif not any(c in t for c in ".eE"):
        t += ".0"
    if ty == "f32":
        t += "f"
    return f"({t})"


def iconst(txt: str) -> str:
    t = txt.replace("_", "")
    v = int(t, 0)
    return f"{v & 0xFFFFFFFF}u"


BIN_I32 = {
    "add": "+", "sub": "-", "mul": "*", "and": "&", "xor": "^", "or": "|",
    "div_u": "/", "rem_u": "%",
}
CMP_I32 = {"eq": "==", "ne": "!=", "lt_u": "<", "le_u": "<=", "gt_u": ">", "ge_u": ">="}
BIN_F = {"add": "+", "sub": "-", "mul": "*", "div": "/"}
CMP_F = {"lt": "<", "gt": ">", "le": "<=", "ge": ">=", "eq": "==", "ne": "!="}

STATEMENT_OPS = {
    "local.set", "global.set", "block", "loop", "if", "br", "br_if", "return", "drop",
    "f32.store", "f64.store", "i32.store", "v128.store", "v128.store64_lane", "nop",
}


class Func:
    def __init__(self, sexpr, index):
        self.params = []   # (cname, type)
        self.locals = []   # (cname, type)
        self.result = None
        self.exports = []
        self.body = []
        items = sexpr[1:]
        self.name = None
        if items and isinstance(items[0], str) and items[0].startswith("$"):
            self.name = items[0]
            items = items[1:]
        else:
            self.name = f"$anon{index}"
        for it in items:
            if isinstance(it, list) and it and it[0] == "export":
                self.exports.append(it[1].strip('"'))
            elif isinstance(it, list) and it and it[0] == "param":
                self._decl(it, self.params)
            elif isinstance(it, list) and it and it[0] == "result":
                self.result = it[1]
            elif isinstance(it, list) and it and it[0] == "local":
                self._decl(it, self.locals)
            else:
                self.body.append(it)
        self.cname = "f_" + cident(self.name)

    @staticmethod
    def _decl(it, dest):
        rest = it[1:]
        if rest and rest[0].startswith("$"):
            dest.append(("l_" + cident(rest[0]), rest[1]))
        else:
            for ty in rest:
                dest.append((f"l_anon{len(dest)}", ty))


class Module:
    def __init__(self, name: str, text: str):
        self.name = name
        top = read_sexprs(text)
        assert len(top) == 1 and top[0][0] == "module"
        self.pages = None
        self.globals = {}        # $name -> (type, c-expression)
        self.global_exports = []  # (export name, type, expr)
        self.funcs = {}
        self.func_list = []
        for item in top[0][1:]:
            kind = item[0]
            if kind == "memory":
                self.pages = int([x for x in item[1:] if isinstance(x, str)][-1])
            elif kind == "global":
                self._global(item)
            elif kind == "func":
                f = Func(item, len(self.func_list))
                self.funcs[f.name] = f
                self.func_list.append(f)
            else:
                raise NotImplementedError(f"module item {kind}")
        self._label_counter = 0

    # ---- globals ------------------------------------------------------
    def _global(self, item):
        rest = item[1:]
        gname = None
        export = None
        if isinstance(rest[0], str) and rest[0].startswith("$"):
            gname = rest[0]
            rest = rest[1:]
        if isinstance(rest[0], list) and rest[0][0] == "export":
            export = rest[0][1].strip('"')
            rest = rest[1:]
        ty = rest[0]
        if isinstance(ty, list):  # (mut T)
            raise NotImplementedError("mutable globals")
        init = self.expr(rest[1], None)
        if gname:
            self.globals[gname] = (ty, init)
        if export:
            self.global_exports.append((export, ty, init))

    # ---- expressions --------------------------------------------------
    def expr(self, e, fn) -> str:
        if isinstance(e, str):
            raise SyntaxError(f"bare atom {e!r} in folded expression")
        op = e[0]
        a = e[1:]
        X = lambda i: self.expr(a[i], fn)  # noqa: E731

        if op == "local.get":
            return self._local(a[0], fn)
        if op == "global.get":
            return f"({self.globals[a[0]][1]})"
        if op == "i32.const":
            return iconst(a[0])
        if op in ("f32.const", "f64.const"):
            return fconst(a[0], op[:3])
        if op == "v128.const":
            shape, vals = a[0], a[1:]
            if shape == "f32x4":
                return "((v128)(f32x4){" + ",".join(fconst(v, "f32") for v in vals) + "})"
            if shape == "f64x2":
                return "((v128)(f64x2){" + ",".join(fconst(v, "f64") for v in vals) + "})"
            if shape == "i32x4":
                return "((v128)(u32x4){" + ",".join(iconst(v) for v in vals) + "})"
            if shape == "i64x2":
                return "((v128)(u64x2){" + ",".join(
                    f"{int(v.replace('_', ''), 0) & 0xFFFFFFFFFFFFFFFF}ull" for v in vals) + "})"
            raise NotImplementedError(shape)
        if op == "call":
            callee = self.funcs[a[0]]
            args = ", ".join(["M"] + [self.expr(x, fn) for x in a[1:]])
            return f"{callee.cname}({args})"

        pre, _, sub = op.partition(".")
        if pre == "i32":
            if sub in BIN_I32:
                return f"((uint32_t)({X(0)} {BIN_I32[sub]} {X(1)}))"
            if sub in CMP_I32:
                return f"((uint32_t)({X(0)} {CMP_I32[sub]} {X(1)}))"
            if sub == "shl":
                return f"((uint32_t)({X(0)} << ({X(1)} & 31u)))"
            if sub == "shr_u":
                return f"((uint32_t)({X(0)} >> ({X(1)} & 31u)))"
            if sub == "eqz":
                return f"((uint32_t)({X(0)} == 0u))"
            if sub == "ctz":
                return f"wat_ctz({X(0)})"
            if sub == "clz":
                return f"wat_clz({X(0)})"
            if sub == "load":
                return f"ld_u32(M, {X(0)})"
        if pre in ("f32", "f64"):
            cty = CTYPE[pre]
            if sub in BIN_F:
                return f"(({cty})({X(0)} {BIN_F[sub]} {X(1)}))"
            if sub in CMP_F:
                return f"((uint32_t)({X(0)} {CMP_F[sub]} {X(1)}))"
            if sub == "neg":
                return f"(-({X(0)}))"
            if sub == "convert_i32_u":
                return f"(({cty})({X(0)}))"
            if sub == "load":
                return f"ld_{pre}(M, {X(0)})"
        if pre in ("f32x4", "f64x2"):
            if sub in ("add", "sub", "mul", "div"):
                return f"((v128)(({pre}){X(0)} {BIN_F[sub]} ({pre}){X(1)}))"
            if sub == "neg":
                return f"((v128)(-({pre}){X(0)}))"
            if sub == "splat":
                return f"{pre}_splat({X(0)})"
            if sub == "extract_lane":
                return f"((({pre}){self.expr(a[1], fn)})[{int(a[0])}])"
            if sub == "replace_lane":
                return f"{pre}_replace({self.expr(a[1], fn)}, {int(a[0])}, {self.expr(a[2], fn)})"
        if op == "i8x16.shuffle":
            idx = [int(x) for x in a[:16]]
            va, vb = self.expr(a[16], fn), self.expr(a[17], fn)
            mask = ",".join(str(i) for i in idx)
            return f"((v128)__builtin_shuffle((u8x16){va}, (u8x16){vb}, (u8x16){{{mask}}}))"
        if op == "v128.load":
            return f"ld_v128(M, {X(0)})"
        if op == "v128.load64_zero":
            return f"ld_v128_64_zero(M, {X(0)})"
        if op == "v128.load64_lane":
            lane = int(a[0])
            return f"ld_v128_64_lane(M, {self.expr(a[1], fn)}, {self.expr(a[2], fn)}, {lane})"
        if op == "v128.xor":
            return f"((v128)({X(0)} ^ {X(1)}))"
        raise NotImplementedError(f"expression op {op}")

    def _local(self, name, fn):
        return "l_" + cident(name)

    # ---- statements ---------------------------------------------------
    def stmts(self, items, fn, labels, ind):
        out = []
        for it in items:
            out.extend(self.stmt(it, fn, labels, ind))
        return out

    def stmt(self, e, fn, labels, ind):
        pad = "  " * ind
        op = e[0]
        a = e[1:]
        if op == "local.set":
            return [f"{pad}{self._local(a[0], fn)} = {self.expr(a[1], fn)};"]
        if op in ("f32.store", "f64.store", "i32.store", "v128.store"):
            ty = {"f32.store": "f32", "f64.store": "f64", "i32.store": "u32", "v128.store": "v128"}[op]
            return [f"{pad}st_{ty}(M, {self.expr(a[0], fn)}, {self.expr(a[1], fn)});"]
        if op == "v128.store64_lane":
            lane = int(a[0])
            return [f"{pad}st_v128_64_lane(M, {self.expr(a[1], fn)}, {self.expr(a[2], fn)}, {lane});"]
        if op == "block" or op == "loop":
            label = None
            body = a
            if body and isinstance(body[0], str) and body[0].startswith("$"):
                label = body[0]
                body = body[1:]
            self._label_counter += 1
            clabel = f"L{self._label_counter}_{cident(label) if label else 'x'}"
            labels = labels + [(label, clabel, op)]
            inner = self.stmts(body, fn, labels, ind + 1)
            if op == "block":
                return [f"{pad}{{"] + inner + [f"{pad}}}", f"{pad}{clabel}: ;"]
            return [f"{pad}{clabel}: ;", f"{pad}{{"] + inner + [f"{pad}}}"]
        if op in ("br", "br_if"):
            target = a[0]
            clabel = None
            if target.startswith("$"):
                for lab, cl, _ in reversed(labels):
                    if lab == target:
                        clabel = cl
                        break
            else:
                clabel = labels[-1 - int(target)][1]
            if clabel is None:
                raise SyntaxError(f"unknown label {target}")
            if op == "br":
                return [f"{pad}goto {clabel};"]
            return [f"{pad}if ({self.expr(a[1], fn)}) goto {clabel};"]
        if op == "if":
            cond = self.expr(a[0], fn)
            then_body, else_body = [], None
            for part in a[1:]:
                if part[0] == "then":
                    then_body = part[1:]
                elif part[0] == "else":
                    else_body = part[1:]
                else:
                    raise NotImplementedError("if without then/else")
            # `if` is itself a br target (depth counting); unnamed
            labels = labels + [(None, None, "if")]
            out = [f"{pad}if ({cond}) {{"] + self.stmts(then_body, fn, labels, ind + 1)
            if else_body is not None:
                out += [f"{pad}}} else {{"] + self.stmts(else_body, fn, labels, ind + 1)
            out.append(f"{pad}}}")
            return out
        if op == "return":
            if a:
                return [f"{pad}return {self.expr(a[0], fn)};"]
            return [f"{pad}return;"]
        if op == "drop":
            return [f"{pad}(void)({self.expr(a[0], fn)});"]
        if op == "nop":
            return []
        if op == "call":
            callee = self.funcs[a[0]]
            if callee.result is None:
                args = ", ".join(["M"] + [self.expr(x, fn) for x in a[1:]])
                return [f"{pad}{callee.cname}({args});"]
        # value expression in statement position: implicit function result
        return [f"{pad}return {self.expr(e, fn)};"]

    # ---- emission -----------------------------------------------------
    def emit(self) -> str:
        out = [f"/* GENERATED by oracle/wat2c.py from modules/{self.name}.wat -- do not commit */",
               '#include "watref_rt.h"', ""]
        for f in self.func_list:
            out.append(self._proto(f) + ";")
        out.append("")
        for f in self.func_list:
            out.append(self._proto(f) + " {")
            for cn, ty in f.locals:
                zero = "{0}" if ty == "v128" else "0"
                out.append(f"  {CTYPE[ty]} {cn} = {zero};")
            out.extend(self.stmts(f.body, f, [], 1))
            out.append("}")
            out.append("")
        out.append(f"uint32_t watref_{self.name}_pages(void) {{ return {self.pages}u; }}")
        for ex, ty, init in self.global_exports:
            out.append(f"{CTYPE[ty]} watref_{self.name}_global_{cident(ex)}(void) {{ return {init}; }}")
        for f in self.func_list:
            for ex in f.exports:
                params = ", ".join(["uint8_t *mem"] + [f"{CTYPE[t]} p{i}" for i, (_, t) in enumerate(f.params)])
                args = ", ".join(["mem"] + [f"p{i}" for i in range(len(f.params))])
                ret = CTYPE[f.result] if f.result else "void"
                call = f"{f.cname}({args})"
                body = f"return {call};" if f.result else f"{call};"
                out.append(f"{ret} watref_{self.name}_{cident(ex)}({params}) {{ {body} }}")
        return "\n".join(out) + "\n"

    def _proto(self, f: Func) -> str:
        params = ", ".join(["uint8_t *M"] + [f"{CTYPE[t]} {cn}" for cn, t in f.params])
        ret = CTYPE[f.result] if f.result else "void"
        return f"static {ret} {f.cname}({params})"


RUNTIME_HEADER = r"""/* GENERATED by oracle/wat2c.py -- runtime shims for transpiled WAT */
#ifndef WATREF_RT_H
#define WATREF_RT_H
#include <stdint.h>
#include <string.h>
#include <math.h>
typedef uint32_t v128  __attribute__((vector_size(16)));
typedef uint32_t u32x4 __attribute__((vector_size(16)));
typedef uint64_t u64x2 __attribute__((vector_size(16)));
typedef uint8_t  u8x16 __attribute__((vector_size(16)));
typedef float    f32x4 __attribute__((vector_size(16)));
typedef double   f64x2 __attribute__((vector_size(16)));
#define WAT_INLINE static inline __attribute__((always_inline, unused))
WAT_INLINE uint32_t wat_ctz(uint32_t x) { return x ? (uint32_t)__builtin_ctz(x) : 32u; }
WAT_INLINE uint32_t wat_clz(uint32_t x) { return x ? (uint32_t)__builtin_clz(x) : 32u; }
WAT_INLINE float    ld_f32(const uint8_t *M, uint32_t a) { float v; memcpy(&v, M + a, 4); return v; }
WAT_INLINE double   ld_f64(const uint8_t *M, uint32_t a) { double v; memcpy(&v, M + a, 8); return v; }
WAT_INLINE uint32_t ld_u32(const uint8_t *M, uint32_t a) { uint32_t v; memcpy(&v, M + a, 4); return v; }
WAT_INLINE v128     ld_v128(const uint8_t *M, uint32_t a) { v128 v; memcpy(&v, M + a, 16); return v; }
WAT_INLINE v128     ld_v128_64_zero(const uint8_t *M, uint32_t a) { u64x2 v = {0, 0}; uint64_t t; memcpy(&t, M + a, 8); v[0] = t; return (v128)v; }
WAT_INLINE v128     ld_v128_64_lane(const uint8_t *M, uint32_t a, v128 x, int lane) { u64x2 v = (u64x2)x; uint64_t t; memcpy(&t, M + a, 8); v[lane] = t; return (v128)v; }
WAT_INLINE void st_f32(uint8_t *M, uint32_t a, float v) { memcpy(M + a, &v, 4); }
WAT_INLINE void st_f64(uint8_t *M, uint32_t a, double v) { memcpy(M + a, &v, 8); }
WAT_INLINE void st_u32(uint8_t *M, uint32_t a, uint32_t v) { memcpy(M + a, &v, 4); }
WAT_INLINE void st_v128(uint8_t *M, uint32_t a, v128 v) { memcpy(M + a, &v, 16); }
WAT_INLINE void st_v128_64_lane(uint8_t *M, uint32_t a, v128 x, int lane) { uint64_t t = ((u64x2)x)[lane]; memcpy(M + a, &t, 8); }
WAT_INLINE v128 f32x4_splat(float x) { return (v128)(f32x4){x, x, x, x}; }
WAT_INLINE v128 f64x2_splat(double x) { return (v128)(f64x2){x, x}; }
WAT_INLINE v128 f32x4_replace(v128 a, int k, float x) { f32x4 t = (f32x4)a; t[k] = x; return (v128)t; }
WAT_INLINE v128 f64x2_replace(v128 a, int k, double x) { f64x2 t = (f64x2)a; t[k] = x; return (v128)t; }
#endif
"""


def main(argv):
    if len(argv) < 3:
        print("usage: wat2c.py <modules-dir> <out-dir>", file=sys.stderr)
        return 2
    src, dst = Path(argv[1]), Path(argv[2])
    dst.mkdir(parents=True, exist_ok=True)
    (dst / "watref_rt.h").write_text(RUNTIME_HEADER)
    for wat in sorted(src.glob("*.wat")):
        mod = Module(wat.stem, wat.read_text())
        (dst / f"{wat.stem}.c").write_text(mod.emit())
        print(f"wat2c: {wat.name} -> {dst / (wat.stem + '.c')}  "
              f"({len(mod.func_list)} funcs, {mod.pages} pages)")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))

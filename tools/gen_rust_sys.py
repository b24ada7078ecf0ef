#!/usr/bin/env python
"""Generate rust/hpt-b200-sys/src/lib.rs from include/hpt_b200.h — ONE source of truth for the C ABI.

    python tools/gen_rust_sys.py            # rewrite the crate's lib.rs
    python tools/gen_rust_sys.py --check    # exit 1 if the committed file is stale

The header is plain C with a regular shape (typedef enum / typedef struct / prototypes), so a small parser is
enough; `parse_header()` is also what tests/test_host.py uses to compare enum VALUES and struct FIELD LISTS with the
ctypes mirror (hpt_b200/_ffi.py).  Enums become `pub type X = c_int` + `pub const`s (an out-of-range value coming back
over FFI must not be undefined behaviour, which it would be for a Rust `enum`)."""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hpt_b200.h")
OUT = os.path.join(ROOT, "rust", "hpt-b200-sys", "src", "lib.rs")

SCALARS = {"int": "c_int", "int32_t": "i32", "int64_t": "i64", "uint8_t": "u8", "uint32_t": "u32", "uint64_t": "u64",
           "size_t": "usize", "double": "c_double", "char": "c_char", "void": "c_void"}


def strip_comments(text):
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def parse_header(path=HEADER):
    """→ dict(defines={name: int}, enums={name: [(member, value)]}, structs={name: [(field, ctype, dims)]},
    opaque=[names], functions=[(name, ret, [(ctype, argname)])])"""
    raw = open(path).read()
    text = strip_comments(raw)
    defines = {m.group(1): int(m.group(2)) for m in re.finditer(r"^#define\s+(HPTB_\w+)\s+(\d+)\s*$", text, flags=re.M)}
    enums, structs, opaque, functions = {}, {}, [], []
    for m in re.finditer(r"typedef\s+enum\s*(\w*)\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        members, nxt = [], 0
        for item in m.group(2).split(","):
            item = item.strip()
            if not item:
                continue
            if "=" in item:
                name, val = [s.strip() for s in item.split("=")]
                nxt = defines[val] if val in defines else int(val, 0)
            else:
                name = item
            members.append((name, nxt))
            nxt += 1
        enums[m.group(3)] = members
    for m in re.finditer(r"typedef\s+struct\s*(\w*)\s*\{(.*?)\}\s*(\w+)\s*;", text, flags=re.S):
        fields = []
        for decl in m.group(2).split(";"):
            decl = " ".join(decl.split())
            if not decl:
                continue
            mm = re.match(r"((?:const\s+)?\w+\s*\**)\s*(.*)$", decl)
            ctype, names = mm.group(1).replace(" *", "*").strip(), mm.group(2)
            for nm in names.split(","):
                nm = nm.strip()
                dims = [d for d in re.findall(r"\[(\w+)\]", nm)]
                fields.append((re.sub(r"\[.*", "", nm), ctype, [defines.get(d, None) if not d.isdigit() else int(d) for d in dims]))
        structs[m.group(3)] = fields
    for m in re.finditer(r"typedef\s+struct\s+(\w+)\s+(\w+)\s*;", text):
        opaque.append(m.group(2))
    body = re.sub(r"typedef\s+(enum|struct)\s*\w*\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)
    for m in re.finditer(r"^([A-Za-z_][\w \*]*?)\s*\b(hptb_\w+)\s*\(([^;{}]*?)\)\s*;", body, flags=re.M | re.S):
        ret, name, args = " ".join(m.group(1).split()), m.group(2), " ".join(m.group(3).split())
        if ret.startswith("typedef") or ret.startswith("#"):
            continue
        arglist = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.+?)(\w+)$", a)
                ctype, an = mm.group(1).strip(), mm.group(2)
                if an in SCALARS or an.startswith("hptb_"):  # unnamed parameter
                    ctype, an = a, "arg%d" % len(arglist)
                arglist.append((" ".join(ctype.split()), an))
        functions.append((name, ret, arglist))
    return {"defines": defines, "enums": enums, "structs": structs, "opaque": opaque, "functions": functions}


def rust_type(ctype, known):
    """C type (possibly with const / pointers) → Rust FFI type."""
    t = ctype.replace("*", " * ").split()
    # parse right-to-left pointer chain: base [const] (* [const])*
    base, i, base_const = None, 0, False
    while i < len(t) and t[i] != "*":
        if t[i] == "const":
            base_const = True
        elif t[i] != "struct":
            base = t[i]
        i += 1
    if base in SCALARS:
        r = SCALARS[base]
    elif base in known:
        r = base if base not in ("hptb_status",) else "hptb_status"
    else:
        raise ValueError(f"unknown C type {ctype!r}")
    const = base_const
    while i < len(t):
        assert t[i] == "*"
        i += 1
        r = ("*const " if const else "*mut ") + r
        const = False
        if i < len(t) and t[i] == "const":
            const = True
            i += 1
    return r


RUST_KEYWORDS = {"in", "type", "ref", "box", "move", "fn", "mod", "use", "loop", "match", "impl", "self", "super", "where", "as",
                 "dyn", "final", "override", "priv", "abstract", "become", "do", "macro", "typeof", "unsized", "virtual", "yield", "try"}


def rust_ident(name):
    return "r#" + name if name in RUST_KEYWORDS else name


def generate():
    h = parse_header()
    known = set(h["enums"]) | set(h["structs"]) | set(h["opaque"])
    o = []
    o.append("//! Raw bindings to include/hpt_b200.h — GENERATED by tools/gen_rust_sys.py, do not edit (tests/test_host.py checks that")
    o.append("//! this file is what the generator produces from the current header, and that the ctypes mirror agrees with both).")
    o.append("//! UNVERIFIED BY A COMPILER: the build image has no Rust toolchain; the same symbols are bound and exercised from")
    o.append("//! Python (hpt_b200/_ffi.py).  Enums are `c_int` aliases with constants: a value outside the list coming back over")
    o.append("//! FFI must not be undefined behaviour.")
    o.append("#![allow(non_camel_case_types, non_upper_case_globals, dead_code)]")
    o.append("use std::os::raw::{c_char, c_double, c_int, c_void};")
    o.append("")
    for k, v in h["defines"].items():
        o.append(f"pub const {k}: usize = {v};")
    o.append("")
    for name, members in h["enums"].items():
        o.append(f"pub type {name} = c_int;")
        for mn, mv in members:
            o.append(f"pub const {mn}: {name} = {mv};")
        o.append("")
    for name in h["opaque"]:
        o.append("#[repr(C)]")
        o.append(f"pub struct {name} {{ _private: [u8; 0] }}")
    o.append("")
    for name, fields in h["structs"].items():
        o.append("#[repr(C)]")
        o.append("#[derive(Clone, Copy)]")
        o.append(f"pub struct {name} {{")
        for fn_, ct, dims in fields:
            rt = rust_type(ct, known)
            for d in reversed(dims):
                rt = f"[{rt}; {d}]"
            o.append(f"    pub {rust_ident(fn_)}: {rt},")
        o.append("}")
        o.append("")
    o.append('#[link(name = "hpt_b200")]')
    o.append('extern "C" {')
    for name, ret, args in h["functions"]:
        al = ", ".join(f"{rust_ident(an)}: {rust_type(ct, known)}" for ct, an in args)
        r = "" if ret == "void" else f" -> {rust_type(ret, known)}"
        o.append(f"    pub fn {name}({al}){r};")
    o.append("}")
    o.append("")
    return "\n".join(o)


def main():
    text = generate()
    if "--check" in sys.argv:
        cur = open(OUT).read() if os.path.exists(OUT) else ""
        if cur != text:
            print("rust/hpt-b200-sys/src/lib.rs is stale: run python tools/gen_rust_sys.py")
            sys.exit(1)
        return
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    with open(OUT, "w") as f:
        f.write(text)
    print("wrote", OUT)


if __name__ == "__main__":
    main()

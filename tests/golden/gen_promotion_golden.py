"""Generate tests/golden/promotion.json from the reference's Rust promotion tables.

Run in the build container only (needs /root/reference):
    python tests/golden/gen_promotion_golden.py

Parses hpt-types/src/promotion/normal_promote/_<T>.rs: every
impl_normal_out_promote!(L, R, Out, Intermediate), impl_float_out_binary_promote!(…)
and impl_float_out_unary_promote!(T, Out, Intermediate) line for the 13 scalar dtypes of the
hot path (isize/usize/complex rows are outside the path).  The JSON is the golden vector the
C++ tables in hpt_b200/csrc/promote.cpp and oracle/hpt_oracle.py are pinned against.
"""
import json, os, re, sys

REF = "/root/reference/hpt-types/src/promotion/normal_promote"
DTYPES = ["bool", "i8", "i16", "i32", "i64", "u8", "u16", "u32", "u64", "f16", "bf16", "f32", "f64"]


def main():
    out = {"dtypes": DTYPES, "normal_out": {}, "float_out_binary": {}, "float_out_unary": {},
           "normal_out_intermediate": {}, "float_out_binary_intermediate": {},
           "float_out_unary_intermediate": {},
           "source": "hpt-types/src/promotion/normal_promote/_*.rs (Hpt v0.1.3)"}
    pat2 = re.compile(r"^(?:\s*)impl_(normal_out|float_out_binary)_promote!\((\w+),\s*(\w+),\s*(\w+),\s*(\w+)\);")
    pat1 = re.compile(r"^(?:\s*)impl_float_out_unary_promote!\((\w+),\s*(\w+),\s*(\w+)\);")
    for l in DTYPES:
        lines = open(os.path.join(REF, f"_{l}.rs")).read().splitlines()
        skip = False
        for i, line in enumerate(lines):
            # skip the 32-bit pointer-width variants (isize/usize rows anyway)
            if 'target_pointer_width = "32"' in line:
                skip = True
                continue
            if skip:
                skip = False
                continue
            m = pat2.match(line)
            if m:
                kind, a, b, o, im = m.groups()
                if a != l or b not in DTYPES:
                    continue
                out[kind].setdefault(a, {})[b] = o
                out[kind + "_intermediate"].setdefault(a, {})[b] = im
                continue
            m = pat1.match(line)
            if m:
                a, o, im = m.groups()
                if a == l:
                    out["float_out_unary"][a] = o
                    out["float_out_unary_intermediate"][a] = im
    for kind in ("normal_out", "float_out_binary"):
        for a in DTYPES:
            assert sorted(out[kind][a]) == sorted(DTYPES), (kind, a, out[kind].get(a))
    assert sorted(out["float_out_unary"]) == sorted(DTYPES)
    dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "promotion.json")
    json.dump(out, open(dst, "w"), indent=1, sort_keys=True)
    print("wrote", dst)


if __name__ == "__main__":
    main()

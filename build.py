#!/usr/bin/env python
"""Build libhpt_b200.so (CUDA kernels + C ABI) and the oracle/baseline helpers, in-tree.

    python build.py            # incremental (ninja)
    python build.py --clean

Every .cu is compiled for sm_100a only (`-gencode arch=compute_100a,code=sm_100a -lineinfo`).
The elementwise kernels are instantiated per op (specialised same-dtype kernels) and per output dtype
(runtime-typed kernels) from template sources with -D flags, so the ~70 units compile in parallel.  Outputs:
    hpt_b200/lib/libhpt_b200.so     the product
    oracle/_build/liboracle_cpu.so  C++/OpenMP restatement of Hpt's CPU path (test oracle + CPU baseline)
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "hpt_b200", "csrc")
BUILD = os.path.join(ROOT, "build")
LIBDIR = os.path.join(ROOT, "hpt_b200", "lib")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

DTYPES = [("b8", "bool"), ("int8_t", "i8"), ("int16_t", "i16"), ("int32_t", "i32"), ("int64_t", "i64"),
          ("uint8_t", "u8"), ("uint16_t", "u16"), ("uint32_t", "u32"), ("uint64_t", "u64"),
          ("f16", "f16"), ("bf16", "bf16"), ("float", "f32"), ("double", "f64")]
BINARY_OPS = [("OpAdd", "add", 0, 1), ("OpSub", "sub", 0, 0), ("OpMul", "mul", 0, 1), ("OpRem", "rem", 0, 0),
              ("OpDiv", "div", 1, 0), ("OpMax", "maximum", 0, 1), ("OpMin", "minimum", 0, 1),
              ("OpPow", "pow", 1, 0), ("OpHypot", "hypot", 1, 0)]
NORMAL_UNARY_OPS = ["floor", "ceil", "round", "trunc", "abs", "neg", "sign", "square", "relu", "relu6", "leaky_relu",
                    "clamp", "bitnot"]
UNARY_OPS = ["sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh",
             "exp", "exp2", "exp10", "ln", "log2", "log10", "sqrt", "cbrt", "recip", "erf", "sigmoid", "gelu",
             "selu", "elu", "celu", "mish", "softplus", "softsign", "hard_sigmoid", "hard_swish"]
REDUCE_OPS = ["sum", "mean", "max", "min", "argmax", "argmin", "logsumexp", "sum_square", "prod",
              "reducel1", "nansum", "nanprod", "all", "any", "reducel2", "reducel3"]

NVCC_FLAGS = ("-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --expt-relaxed-constexpr "
              "-Xfatbin -compress-all -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -Xcudafe --diag_suppress=177 "
              "-Xcudafe --diag_suppress=550 -I%s" % os.path.join(ROOT, "include"))
CXX_FLAGS = "-O2 -std=c++17 -fPIC -fvisibility=hidden -I/usr/local/cuda/include -I%s" % os.path.join(ROOT, "include")


def units():
    """(object name, source, extra defines)"""
    u = []
    # specialised vector-only kernels: same-dtype binary, float unary, same-dtype copy
    for fn, name, kind, bool_ok in BINARY_OPS:
        u.append((f"binary_{name}", "binary_inst.cu",
                  f"-DHPTB_OP={fn} -DHPTB_OPNAME={name} -DHPTB_KIND={kind} -DHPTB_BOOL_OK={bool_ok}"))
    for name in UNARY_OPS:
        u.append((f"unary_{name}", "unary_inst.cu", f"-DHPTB_OPENUM=HPTB_{name.upper()} -DHPTB_OPNAME={name}"))
    for name in NORMAL_UNARY_OPS:
        u.append((f"nunary_{name}", "nunary_inst.cu", f"-DHPTB_OPENUM=HPTB_{name.upper()} -DHPTB_OPNAME={name}"))
    u.append(("copy_same", "cast_inst.cu", ""))
    # runtime-typed kernels: one unit per OUTPUT dtype (mixed-dtype binary, integer-input unary, astype)
    for cty, short in DTYPES:
        is_float = 1 if short in ("f16", "bf16", "f32", "f64") else 0
        u.append((f"dyn_{short}", "dyn_inst.cu", f"-DHPTB_OUT={cty} -DHPTB_OUTNAME={short} -DHPTB_OUT_FLOAT={is_float}"))
    for name in REDUCE_OPS:
        u.append((f"reduce_{name}", "reduce_inst.cu", f"-DHPTB_OPENUM=HPTB_{name.upper()} -DHPTB_OPNAME={name}"
                  + (" -DHPTB_LOGSUMEXP_LONG=1" if name == "logsumexp" else "")))
    for fn, name in (("OpAdd", "add"), ("OpSub", "sub"), ("OpMul", "mul")):
        u.append((f"fused_{name}", "fused_inst.cu", f"-DHPTB_OP={fn} -DHPTB_OPNAME={name}"))
    for src in ("softmax.cu", "misc.cu", "meanvar.cu", "sharded.cu"):
        if os.path.exists(os.path.join(CSRC, src)):
            u.append((src[:-3], src, ""))
    return u


def write_ninja(variant="", extra="", only=()):
    """`only`: unit-name prefixes that get the variant's flags; the other units reuse the default build's objects."""
    global BUILD
    default_obj = os.path.join(BUILD, "obj")
    if variant:
        BUILD = os.path.join(ROOT, "build", "variant_" + variant)
    os.makedirs(BUILD, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    lines = [
        "ninja_required_version = 1.3",
        f"nvcc = {NVCC}",
        f"nvflags = {NVCC_FLAGS} {extra}",
        f"cxxflags = {CXX_FLAGS}",
        "rule nvcc",
        "  command = $nvcc $nvflags $defs -MD -MF $out.d -c $in -o $out",
        "  depfile = $out.d",
        "  deps = gcc",
        "  description = NVCC $out",
        "rule cxx",
        "  command = g++ $cxxflags -MD -MF $out.d -c $in -o $out",
        "  depfile = $out.d",
        "  deps = gcc",
        "  description = CXX $out",
        "rule link",
        "  command = g++ -shared -o $out @$out.rsp -L/usr/local/cuda/lib64 -lcudart_static -Wl,--exclude-libs,ALL -ldl -lrt -lpthread",
        "  rspfile = $out.rsp",
        "  rspfile_content = $in",
        "  description = LINK $out",
        "",
    ]
    objs = []
    for name, src, defs in units():
        if variant and only and not any(name.startswith(o) for o in only):
            objs.append(os.path.join(default_obj, name + ".o"))  # built by the default configuration
            continue
        obj = os.path.join(BUILD, "obj", name + ".o")
        lines += [f"build {obj}: nvcc {os.path.join(CSRC, src)}", f"  defs = {defs}"]
        objs.append(obj)
    for src in sorted(os.listdir(CSRC)):
        if src.endswith(".cpp"):
            if variant and only:
                objs.append(os.path.join(default_obj, src[:-4] + ".o"))
                continue
            obj = os.path.join(BUILD, "obj", src[:-4] + ".o")
            lines += [f"build {obj}: cxx {os.path.join(CSRC, src)}"]
            objs.append(obj)
    lib = os.path.join(LIBDIR, f"libhpt_b200_{variant}.so" if variant else "libhpt_b200.so")
    lines += [f"build {lib}: link {' '.join(objs)}", f"default {lib}", ""]
    with open(os.path.join(BUILD, "build.ninja"), "w") as f:
        f.write("\n".join(lines))
    return lib


def build_oracle():
    """C++/OpenMP restatement of Hpt's CPU path: the parity oracle and the timed CPU baseline."""
    src = os.path.join(ROOT, "oracle", "oracle_cpu.cpp")
    if not os.path.exists(src):
        return None
    outdir = os.path.join(ROOT, "oracle", "_build")
    os.makedirs(outdir, exist_ok=True)
    out = os.path.join(outdir, "liboracle_cpu.so")
    if os.path.exists(out) and os.path.getmtime(out) >= os.path.getmtime(src):
        return out
    # x86-64-v3 (AVX2+FMA) rather than -march=native: the .so travels to the GPU box, whose host
    # CPU may differ from the build container's.
    cmd = ["g++", "-O3", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-march=x86-64-v3", "-ffp-contract=off",
           "-fno-math-errno", "-o", out, src, "-lmvec", "-lm"]
    subprocess.check_call(cmd)
    return out


def main():
    if "--clean" in sys.argv:
        shutil.rmtree(BUILD, ignore_errors=True)
        shutil.rmtree(LIBDIR, ignore_errors=True)
        shutil.rmtree(os.path.join(ROOT, "oracle", "_build"), ignore_errors=True)
    # tuning builds: `python build.py --variant u8 -DHPTB_RED_UNROLL=8 …` → hpt_b200/lib/libhpt_b200_u8.so
    # `--only reduce,meanvar` limits the variant's flags to those units (the rest is taken from the default build)
    variant, extra, only = "", [], ()
    args = sys.argv[1:]
    if "--variant" in args:
        variant = args[args.index("--variant") + 1]
        extra = [a for a in args if a.startswith("-D")]
        if "--only" in args:
            only = tuple(args[args.index("--only") + 1].split(","))
    lib = write_ninja(variant, " ".join(extra), only)
    jobs = os.environ.get("HPTB_BUILD_JOBS", str(os.cpu_count() or 4))
    subprocess.check_call(["ninja", "-C", BUILD, "-j", jobs] + (["-v"] if "-v" in sys.argv else []))
    if not variant:
        build_oracle()
    print("built", lib)


if __name__ == "__main__":
    main()

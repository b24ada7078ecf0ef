"""Loader for the REFERENCE's own CUDA kernels (oracle/_ref/*.cubin, built by oracle/build_ref.sh from
/root/reference/hpt-cudakernels/src where the sources lie): CUDA driver API through ctypes, launched the way the
reference's Rust host code launches them.  Test infrastructure only.

Launch shapes restate the reference's host logic:
  contiguous_{argmax,argmin}_<T>(out, buffer, in, finished, size)      fast_all_reduce, hpt/src/backends/cuda/utils/reduce/reduce.rs:231-283
  contiguous_{op}_small_fast_dim_only_<T>(out, in, fast_dim, outputs)  reduce.rs:403-437 (case 2) + arg_template.cuh:89-111
  strided_copy_<T>(dst, src, FastDivmod* shape, i32* strides, ndim, n) hpt-cudakernels/src/strided_copy.cu:6-21
  <op>_<L>_<R>_contiguous(out, lhs, rhs, i32 n)                        binary_template.cuh:104-108, launched by binary_fn_precompiled
  <op>_<L>_<R>_uncontiguous(out, lhs, FastDivmod* lshape, i32* lstrides, rhs, FastDivmod* rshape, i32* rstrides,
                            i32 lndim, i32 rndim, i32 n)               (hpt/src/backends/cuda/utils/binary/binary_normal.rs:371-544)
    — one element per thread: the kernels' grid-stride loops index with the thread id, not the loop variable
      (binary_template.cuh:9-13), so the reference is only right when the grid covers n; the launches here do
FastDivmod{i32 divisor; u32 multiplier; u32 shift_right} is CUTLASS's find_divisor as restated in
hpt/src/backends/common/divmod.rs:1-52.
"""
import ctypes
import os
import struct

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "oracle", "_ref")


def available():
    return all(os.path.exists(os.path.join(REF_DIR, n + ".cubin")) for n in ("argmax", "argmin", "strided_copy"))


def available_binary(names=("add", "sub", "mul", "rem")):
    return all(os.path.exists(os.path.join(REF_DIR, f"binary_{n}.cubin")) for n in names)


class RefModule:
    _cuda = None

    def __init__(self, name):
        if RefModule._cuda is None:
            RefModule._cuda = ctypes.CDLL("libcuda.so.1")
        self.cu = RefModule._cuda
        torch.cuda.init()
        torch.zeros(1, device="cuda")  # the primary context is current on this thread
        self.mod = ctypes.c_void_p()
        rc = self.cu.cuModuleLoad(ctypes.byref(self.mod), os.path.join(REF_DIR, name + ".cubin").encode())
        assert rc == 0, f"cuModuleLoad({name}) failed: {rc}"
        self.fns = {}

    def has(self, name):
        f = ctypes.c_void_p()
        return self.cu.cuModuleGetFunction(ctypes.byref(f), self.mod, name.encode()) == 0

    def fn(self, name):
        if name not in self.fns:
            f = ctypes.c_void_p()
            rc = self.cu.cuModuleGetFunction(ctypes.byref(f), self.mod, name.encode())
            assert rc == 0, f"cuModuleGetFunction({name}) failed: {rc}"
            self.fns[name] = f
        return self.fns[name]

    def launch(self, name, grid, block, args):
        """args: list of ctypes values (c_void_p for device pointers, c_size_t / c_int64 / c_int32 for scalars)"""
        ptrs = (ctypes.c_void_p * len(args))(*[ctypes.cast(ctypes.pointer(a), ctypes.c_void_p) for a in args])
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        rc = self.cu.cuLaunchKernel(self.fn(name), grid[0], grid[1], grid[2], block[0], block[1], block[2], 0, stream, ptrs, None)
        assert rc == 0, f"cuLaunchKernel({name}) failed: {rc}"
        torch.cuda.synchronize()


def fast_divmod(d):
    """hpt/src/backends/common/divmod.rs:31-52 (CUTLASS find_divisor): 12-byte {divisor, multiplier, shift_right}"""
    if d == 1:
        return struct.pack("<iII", 1, 0, 0)
    log2 = (d - 1).bit_length()  # ceil(log2 d)
    p = 31 + log2
    m = ((1 << p) + d - 1) // d
    return struct.pack("<iII", d, m & 0xFFFFFFFF, p - 32)


def ref_arg_flat(mod, op, x):
    """full argmax/argmin of a contiguous f32 tensor through the reference's all-reduce kernel"""
    n = x.numel()
    block = min(512, 1 << (max(32, (n + 31) // 32 * 32).bit_length() - 1))
    grid = max(1, min(148 * 4, (n + block * 4 - 1) // (block * 4)))
    out = torch.zeros(1, dtype=torch.int64, device="cuda")
    buf = torch.zeros(grid * 2, dtype=torch.int64, device="cuda")  # ArgMaxResult<f32> = {f32 val; i64 idx}: 16 bytes
    fin = torch.zeros(1, dtype=torch.int32, device="cuda")
    mod.launch(f"contiguous_{op}_f32", (grid, 1, 1), (block, 1, 1),
               [ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(buf.data_ptr()), ctypes.c_void_p(x.data_ptr()),
                ctypes.c_void_p(fin.data_ptr()), ctypes.c_size_t(n)])
    return out.cpu()


def ref_arg_rows(mod, op, x):
    """argmax/argmin over the last axis of a contiguous [rows, cols] f32 tensor (small_fast_dim_only kernel)"""
    rows, cols = x.shape
    bx = (cols // 32) * 32
    bx = max(32, min(512, bx))
    bx = 1 << (bx.bit_length() - 1)  # last_power_of_two
    grid = max(1, min(rows, 148 * 16))
    out = torch.zeros(rows, dtype=torch.int64, device="cuda")
    mod.launch(f"contiguous_{op}_small_fast_dim_only_f32", (grid, 1, 1), (bx, 1, 1),
               [ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(x.data_ptr()), ctypes.c_size_t(cols), ctypes.c_size_t(rows)])
    return out.cpu()


def ref_strided_copy(mod, view, tname="f32"):
    """contiguous() of a strided torch CUDA view through the reference's strided_copy_<T>"""
    shape, strides = list(view.shape), list(view.stride())
    n = view.numel()
    table = torch.frombuffer(bytearray(b"".join(fast_divmod(s) for s in shape)), dtype=torch.uint8).cuda()
    st = torch.tensor(strides, dtype=torch.int32, device="cuda")
    dst = torch.empty(shape, dtype=view.dtype, device="cuda")
    mod.launch(f"strided_copy_{tname}", (min(148 * 8, (n + 255) // 256), 1, 1), (256, 1, 1),
               [ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(view.data_ptr()), ctypes.c_void_p(table.data_ptr()),
                ctypes.c_void_p(st.data_ptr()), ctypes.c_int32(len(shape)), ctypes.c_int64(n)])
    return dst.cpu()


def ref_binary_contiguous(mod, op, ld, rd, lhs, rhs, out):
    """out[i] = op(lhs[i], rhs[i]) through the reference's <op>_<L>_<R>_contiguous; torch CUDA tensors, out preallocated"""
    n = out.numel()
    mod.launch(f"{op}_{ld}_{rd}_contiguous", ((n + 255) // 256, 1, 1), (256, 1, 1),
               [ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(lhs.data_ptr()), ctypes.c_void_p(rhs.data_ptr()), ctypes.c_int32(n)])


def ref_binary_uncontiguous(mod, op, ld, rd, lhs, rhs, out):
    """lhs / rhs: torch CUDA VIEWS already expanded to out's shape (stride 0 on broadcast dims), as the reference's host
    code passes them"""
    n = out.numel()
    shape = list(out.shape)
    table = torch.frombuffer(bytearray(b"".join(fast_divmod(s) for s in shape)), dtype=torch.uint8).cuda()
    ls = torch.tensor(list(lhs.stride()), dtype=torch.int32, device="cuda")
    rs = torch.tensor(list(rhs.stride()), dtype=torch.int32, device="cuda")
    mod.launch(f"{op}_{ld}_{rd}_uncontiguous", ((n + 255) // 256, 1, 1), (256, 1, 1),
               [ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(lhs.data_ptr()), ctypes.c_void_p(table.data_ptr()),
                ctypes.c_void_p(ls.data_ptr()), ctypes.c_void_p(rhs.data_ptr()), ctypes.c_void_p(table.data_ptr()),
                ctypes.c_void_p(rs.data_ptr()), ctypes.c_int32(len(shape)), ctypes.c_int32(len(shape)), ctypes.c_int32(n)])

//! Replaces the body of `binary_fn_precompiled` (hpt/src/backends/cuda/utils/binary/binary_normal.rs:371-544).
//! What disappears: the four dispatch branches (scalar lhs / scalar rhs with a blocking D2H read of the scalar,
//! :393-483; same-shape contiguous, :484-505; general, :506-543), `to_broadcast_layout`, the four `htod_sync_copy`
//! table uploads and `load_ptx_and_get_data`.  A scalar operand is a stride-0 view; broadcasting is the library's.
use std::borrow::BorrowMut;

use hpt_b200_sys as sys;
use hpt_common::error::{base::TensorError, shape::ShapeError};
use hpt_common::shape::shape_utils::predict_broadcast_shape;

use crate::{as_c, check, ctx, stream, HptbDtype};
// from the `hpt` crate: _Tensor, Cuda, CommonBounds, Allocator, AllocatorOutputRetrive
use hpt::{backend::Cuda, tensor_base::_Tensor};
use hpt_allocator::traits::{Allocator, AllocatorOutputRetrive};
use hpt_traits::tensor::{CommonBounds, TensorInfo};

fn binary_op(op_name: &str) -> Option<(bool, i32)> {
    // (is a comparison, code)
    Some(match op_name {
        "add" => (false, sys::HPTB_ADD), "sub" => (false, sys::HPTB_SUB), "mul" => (false, sys::HPTB_MUL),
        "rem" => (false, sys::HPTB_REM), "div" => (false, sys::HPTB_DIV), "max" => (false, sys::HPTB_MAXIMUM),
        "min" => (false, sys::HPTB_MINIMUM), "pow" => (false, sys::HPTB_POW), "hypot" => (false, sys::HPTB_HYPOT),
        "bitand" => (false, sys::HPTB_BITAND), "bitor" => (false, sys::HPTB_BITOR), "bitxor" => (false, sys::HPTB_BITXOR),
        "shl" => (false, sys::HPTB_SHL), "shr" => (false, sys::HPTB_SHR),
        "eq" => (true, sys::HPTB_EQ), "ne" => (true, sys::HPTB_NE), "lt" => (true, sys::HPTB_LT), "le" => (true, sys::HPTB_LE),
        "gt" => (true, sys::HPTB_GT), "ge" => (true, sys::HPTB_GE),
        _ => return None,
    })
}

/// `extract_out` (binary_normal.rs:546-564), unchanged in meaning: a supplied `out` is validated and ALIASED.
fn extract_out<K, O, const D: usize, Al>(res_shape: &hpt_common::shape::shape::Shape, out: Option<O>) -> Result<_Tensor<K, Cuda, D, Al>, TensorError>
where
    K: CommonBounds + HptbDtype,
    O: BorrowMut<_Tensor<K, Cuda, D, Al>>,
    Al: Allocator,
    Al::Output: AllocatorOutputRetrive,
{
    if let Some(mut out) = out {
        ShapeError::check_inplace_out_layout_valid(res_shape, &out.borrow().layout())?;
        Ok((*out.borrow_mut()).clone())
    } else {
        _Tensor::<K, Cuda, D, Al>::empty(res_shape)
    }
}

#[track_caller]
pub(crate) fn binary_fn_precompiled<A, B, O, K, const CUDA_DEVICE: usize, Al>(
    lhs: &_Tensor<A, Cuda, CUDA_DEVICE, Al>,
    rhs: &_Tensor<B, Cuda, CUDA_DEVICE, Al>,
    op_name: &str,
    _meta: &(), // the phf kernel table of the reference: no longer consulted
    out: Option<O>,
) -> Result<_Tensor<K, Cuda, CUDA_DEVICE, Al>, TensorError>
where
    A: CommonBounds + HptbDtype,
    B: CommonBounds + HptbDtype,
    O: BorrowMut<_Tensor<K, Cuda, CUDA_DEVICE, Al>>,
    K: CommonBounds + HptbDtype,
    Al: Allocator,
    Al::Output: AllocatorOutputRetrive,
{
    let (is_cmp, code) = binary_op(op_name).expect("op_name not found");
    // the promotion tables are the library's (generated from hpt-types/src/promotion/normal_promote/_*.rs): K must agree
    let want = if is_cmp { sys::HPTB_BOOL } else { unsafe { sys::hptb_binary_out_dtype(code, A::HPTB_DTYPE, B::HPTB_DTYPE) } };
    debug_assert_eq!(want, K::HPTB_DTYPE, "output type of {op_name} disagrees with the promotion table");
    let res_shape = predict_broadcast_shape(lhs.shape(), rhs.shape())?;
    let res = extract_out::<K, O, CUDA_DEVICE, Al>(&res_shape, out)?;
    let (l, r) = (as_c(lhs.ptr().ptr, &lhs.layout()), as_c(rhs.ptr().ptr, &rhs.layout()));
    let mut o = as_c(res.ptr().ptr, &res.layout());
    let c = ctx(CUDA_DEVICE)?;
    check(unsafe {
        if is_cmp { sys::hptb_compare(c, code, &l, &r, &mut o, stream()) } else { sys::hptb_binary(c, code, &l, &r, &mut o, stream()) }
    })?;
    Ok(res)
}

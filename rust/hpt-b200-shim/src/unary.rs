//! Replaces the bodies of `uary_fn_precompiled` and `uary_fn_precompiled_1scalar`
//! (hpt/src/backends/cuda/utils/unary/unary.rs:122-196, :199-…): the contiguous / uncontiguous split, the shape and
//! stride table uploads and the PTX lookup go away; any layout is one call.
use std::borrow::BorrowMut;

use hpt_b200_sys as sys;
use hpt_common::error::{base::TensorError, shape::ShapeError};

use crate::{as_c, check, ctx, stream, HptbDtype};
use hpt::{backend::Cuda, tensor_base::_Tensor};
use hpt_allocator::traits::{Allocator, AllocatorOutputRetrive};
use hpt_traits::tensor::{CommonBounds, TensorInfo};

/// the reference's kernel-name stems (hpt-cudakernels/src/unary/*.cu) → hptb_unary_op
fn unary_op(op: &str) -> Option<i32> {
    const NAMES: [&str; 45] = [
        "sin", "cos", "tan", "asin", "acos", "atan", "sinh", "cosh", "tanh", "asinh", "acosh", "atanh", "exp", "exp2", "exp10", "ln",
        "log2", "log10", "sqrt", "cbrt", "recip", "erf", "sigmoid", "gelu", "selu", "elu", "celu", "mish", "softplus", "softsign",
        "hard_sigmoid", "hard_swish", "floor", "ceil", "round", "trunc", "abs", "neg", "sign", "square", "relu", "relu6", "leaky_relu",
        "clamp", "bitnot",
    ];
    NAMES.iter().position(|n| *n == op).map(|i| i as i32) // the list is in hptb_unary_op order (checked by tests/test_host.py)
}

#[track_caller]
pub(crate) fn uary_fn_precompiled<A, O, K, const DEVICE_ID: usize, Al>(
    inp: &_Tensor<A, Cuda, DEVICE_ID, Al>,
    op: &str,
    _meta: &(),
    out: Option<O>,
) -> Result<_Tensor<K, Cuda, DEVICE_ID, Al>, TensorError>
where
    A: CommonBounds + HptbDtype,
    K: CommonBounds + HptbDtype,
    O: BorrowMut<_Tensor<K, Cuda, DEVICE_ID, Al>>,
    Al: Allocator,
    Al::Output: AllocatorOutputRetrive,
{
    uary_fn_precompiled_scalars::<A, O, K, DEVICE_ID, Al>(inp, op, 0.0, 0.0, out)
}

/// `uary_fn_precompiled_1scalar` and the two-scalar ops (selu: alpha, scale; clamp: min, max) in one body.
#[track_caller]
pub(crate) fn uary_fn_precompiled_scalars<A, O, K, const DEVICE_ID: usize, Al>(
    inp: &_Tensor<A, Cuda, DEVICE_ID, Al>,
    op: &str,
    alpha: f64,
    beta: f64,
    out: Option<O>,
) -> Result<_Tensor<K, Cuda, DEVICE_ID, Al>, TensorError>
where
    A: CommonBounds + HptbDtype,
    K: CommonBounds + HptbDtype,
    O: BorrowMut<_Tensor<K, Cuda, DEVICE_ID, Al>>,
    Al: Allocator,
    Al::Output: AllocatorOutputRetrive,
{
    let code = unary_op(op).expect("op not found");
    debug_assert_eq!(unsafe { sys::hptb_unary_out_dtype(code, A::HPTB_DTYPE) }, K::HPTB_DTYPE);
    let ret = if let Some(mut out) = out {
        ShapeError::check_inplace_out_layout_valid(inp.shape(), &out.borrow().layout())?;
        (*out.borrow_mut()).clone()
    } else {
        _Tensor::<K, Cuda, DEVICE_ID, Al>::empty(inp.shape())?
    };
    let i = as_c(inp.ptr().ptr, &inp.layout());
    let mut o = as_c(ret.ptr().ptr, &ret.layout());
    check(unsafe { sys::hptb_unary(ctx(DEVICE_ID)?, code, &i, &mut o, alpha, beta, stream()) })?;
    Ok(ret)
}

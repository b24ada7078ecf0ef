//! Replaces the body of `contiguous_softmax` (hpt/src/backends/cuda/tensor_internal/softmax.rs:85-…): the
//! warp / block / block_large kernel choice, the `uncontiguous` fallbacks and the permute-to-last go away.
use hpt_b200_sys as sys;
use hpt_common::error::{base::TensorError, shape::ShapeError};

use crate::{as_c, check, ctx, stream, HptbDtype};
use hpt::{backend::Cuda, tensor_base::_Tensor};
use hpt_allocator::traits::{Allocator, AllocatorOutputRetrive};
use hpt_traits::tensor::{CommonBounds, TensorInfo};

#[track_caller]
pub(crate) fn contiguous_softmax<T, O, const DEVICE: usize, A>(
    a: &_Tensor<T, Cuda, DEVICE, A>,
    axis: i64,
    c: Option<_Tensor<O, Cuda, DEVICE, A>>,
    is_log_softmax: bool,
) -> Result<_Tensor<O, Cuda, DEVICE, A>, TensorError>
where
    T: CommonBounds + HptbDtype,
    O: CommonBounds + HptbDtype,
    A: Allocator + Send + Sync,
    A::Output: AllocatorOutputRetrive,
{
    let nd = a.ndim() as i64;
    let axis = if axis < 0 { axis + nd } else { axis };
    if axis < 0 || axis >= nd {
        return Err(ShapeError::DimOutOfRange { expected: 0..nd, actual: axis, location: std::panic::Location::caller() }.into());
    }
    let res = if let Some(out) = c {
        ShapeError::check_inplace_out_layout_valid(a.shape(), &out.layout())?;
        out
    } else {
        _Tensor::<O, Cuda, DEVICE, A>::empty(a.shape())?
    };
    let i = as_c(a.ptr().ptr, &a.layout());
    let mut o = as_c(res.ptr().ptr, &res.layout());
    check(unsafe { sys::hptb_softmax(ctx(DEVICE)?, &i, axis as i32, is_log_softmax as i32, &mut o, stream()) })?;
    Ok(res)
}

//! Replaces `reduce`, `reduce2`, `reduce3` (hpt/src/backends/cuda/utils/reduce/reduce.rs:62-217) and everything they
//! call: `contiguous_reduce` / `uncontiguous_reduce` / `fast_all_reduce` / `not_keep_last_dim` / `keep_last_dim`
//! (:231-838), `contiguous_reduce_template` (reduce_template.rs:17-82), `reduce_prepare` (reduce_utils.rs:18-74).
//! No per-call `cuMemAlloc` of scratch and ticket buffers, no `set_val` launch for `init_val` (the kernel starts from
//! the op's identity; `init_out = false` folds the previous contents of `c` in), no permute-to-last + `contiguous()`.
use hpt_b200_sys as sys;
use hpt_common::error::{base::TensorError, shape::ShapeError};

use crate::{as_c, check, ctx, stream, HptbDtype};
use hpt::{backend::Cuda, tensor_base::_Tensor};
use hpt_allocator::traits::{Allocator, AllocatorOutputRetrive};
use hpt_traits::tensor::{CommonBounds, TensorInfo};

/// the reference's `op` strings (hpt/src/backends/cuda/tensor_internal/{common_reduce,arg_reduce}.rs) → hptb_reduce_op
fn reduce_op(op: &str) -> Option<i32> {
    Some(match op {
        "sum" => sys::HPTB_SUM, "mean" => sys::HPTB_MEAN, "max" => sys::HPTB_MAX, "min" => sys::HPTB_MIN,
        "argmax" => sys::HPTB_ARGMAX, "argmin" => sys::HPTB_ARGMIN, "logsumexp" => sys::HPTB_LOGSUMEXP,
        "sum_square" => sys::HPTB_SUM_SQUARE, "prod" => sys::HPTB_PROD, "reducel1" => sys::HPTB_REDUCEL1,
        "nansum" => sys::HPTB_NANSUM, "nanprod" => sys::HPTB_NANPROD, "all" => sys::HPTB_ALL, "any" => sys::HPTB_ANY,
        "reducel2" => sys::HPTB_REDUCEL2, "reducel3" => sys::HPTB_REDUCEL3,
        _ => return None,
    })
}

/// The one body behind the three entry points.  `axes` are already normalised by `process_axes` in the trait impl
/// (hpt-common/src/axis/axis.rs:38-72), as today.
#[track_caller]
fn reduce_any<T, O, const DEVICE_ID: usize, Al>(
    a: &_Tensor<T, Cuda, DEVICE_ID, Al>,
    axes: &[usize],
    keepdims: bool,
    init_out: bool,
    op: &str,
    c: Option<_Tensor<O, Cuda, DEVICE_ID, Al>>,
) -> Result<_Tensor<O, Cuda, DEVICE_ID, Al>, TensorError>
where
    T: CommonBounds + HptbDtype,
    O: CommonBounds + HptbDtype,
    Al: Allocator,
    Al::Output: AllocatorOutputRetrive,
{
    let code = reduce_op(op).expect("op not found");
    debug_assert_eq!(unsafe { sys::hptb_reduce_out_dtype(code, T::HPTB_DTYPE) }, O::HPTB_DTYPE);
    // Layout::reduce (hpt-common/src/layout/layout_utils.rs:310-349): keep_dims = false shape, [1] when all axes go
    let res_layout = a.layout().reduce(axes, false)?;
    let res = if let Some(out) = c {
        // reduce_prepare's rule (reduce_utils.rs:30-45): the supplied buffer must hold exactly the reduced size
        ShapeError::check_inplace_out_layout_valid(res_layout.shape(), &out.layout())?;
        out
    } else {
        _Tensor::<O, Cuda, DEVICE_ID, Al>::empty(res_layout.shape())?
    };
    let ax: Vec<i32> = axes.iter().map(|v| *v as i32).collect();
    let i = as_c(a.ptr().ptr, &a.layout());
    let mut o = as_c(res.ptr().ptr, &res_layout); // the keep_dims = false view of `res`
    check(unsafe { sys::hptb_reduce(ctx(DEVICE_ID)?, code, &i, ax.as_ptr(), ax.len() as i32, &mut o, init_out as i32, stream()) })?;
    if keepdims {
        // the same reshape the reference ends with (reduce.rs:837)
        res.reshape(a.layout().reduce(axes, true)?.shape())
    } else {
        Ok(res)
    }
}

#[track_caller]
pub(crate) fn reduce<T, BufferType, const DEVICE_ID: usize, Al>(
    a: &_Tensor<T, Cuda, DEVICE_ID, Al>, axes: &[usize], _init_val: T, keepdims: bool, init_out: bool, _meta: &(), _module_name: &str,
    op: &str, c: Option<_Tensor<T, Cuda, DEVICE_ID, Al>>,
) -> Result<_Tensor<T, Cuda, DEVICE_ID, Al>, TensorError>
where T: CommonBounds + HptbDtype, Al: Allocator, Al::Output: AllocatorOutputRetrive {
    reduce_any::<T, T, DEVICE_ID, Al>(a, axes, keepdims, init_out, op, c)
}

#[track_caller]
pub(crate) fn reduce2<T, O, BufferType, const DEVICE_ID: usize, Al>(
    a: &_Tensor<T, Cuda, DEVICE_ID, Al>, axes: &[usize], _init_val: O, keepdims: bool, init_out: bool, _meta: &(), _module_name: &str,
    op: &str, c: Option<_Tensor<O, Cuda, DEVICE_ID, Al>>,
) -> Result<_Tensor<O, Cuda, DEVICE_ID, Al>, TensorError>
where T: CommonBounds + HptbDtype, O: CommonBounds + HptbDtype, Al: Allocator, Al::Output: AllocatorOutputRetrive {
    reduce_any::<T, O, DEVICE_ID, Al>(a, axes, keepdims, init_out, op, c)
}

/// `reduce3` carried a post-op closure on the host side in the CPU backend and a `_post` kernel on CUDA (mean: ÷ n,
/// logsumexp: ln); both are inside the library's kernel now.
#[track_caller]
pub(crate) fn reduce3<T, O, BufferType, const DEVICE_ID: usize, Al>(
    a: &_Tensor<T, Cuda, DEVICE_ID, Al>, axes: &[usize], _init_val: O, keepdims: bool, init_out: bool, _meta: &(), _module_name: &str,
    op: &str, c: Option<_Tensor<O, Cuda, DEVICE_ID, Al>>,
) -> Result<_Tensor<O, Cuda, DEVICE_ID, Al>, TensorError>
where T: CommonBounds + HptbDtype, O: CommonBounds + HptbDtype, Al: Allocator, Al::Output: AllocatorOutputRetrive {
    reduce_any::<T, O, DEVICE_ID, Al>(a, axes, keepdims, init_out, op, c)
}

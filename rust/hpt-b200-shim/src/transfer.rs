//! Replaces the bodies of `to_cuda::<N>()` (hpt/src/backends/cpu/tensor_impls.rs:295-313), `to_cpu::<N>()`
//! (hpt/src/backends/cuda/tensor_impls.rs:142-169) and `Contiguous::contiguous`
//! (hpt/src/backends/cuda/tensor_internal/normal_out_unary.rs:280-309).  cudarc's `htod_sync_copy_into` /
//! `upgrade_device_ptr` / `leak` dance is one `hptb_memcpy_h2d`; a view is gathered on the device by `hptb_copy`
//! (strided_copy.cu's replacement) instead of on the host.
use std::os::raw::c_void;

use hpt_b200_sys as sys;
use hpt_common::error::base::TensorError;

use crate::{as_c, check, ctx, stream, HptbDtype};
use hpt::{backend::{Cpu, Cuda}, tensor_base::_Tensor};
use hpt_allocator::traits::{Allocator, AllocatorOutputRetrive};
use hpt_traits::tensor::{CommonBounds, TensorInfo};

pub fn to_cuda<T, const CUDA_DEVICE: usize, A>(
    host: &_Tensor<T, Cpu, 0, A>,
) -> Result<_Tensor<T, Cuda, CUDA_DEVICE, <A as Allocator>::CudaAllocator>, TensorError>
where
    T: CommonBounds + HptbDtype,
    A: Allocator,
    A::Output: AllocatorOutputRetrive,
    <<A as Allocator>::CudaAllocator as Allocator>::Output: AllocatorOutputRetrive,
{
    let data = _Tensor::<T, Cuda, CUDA_DEVICE, <A as Allocator>::CudaAllocator>::empty(host.shape())?;
    // a host view is gathered on the host first, as the reference does (`self.contiguous()?`, :306-309)
    let dense;
    let src = if host.is_contiguous() && host.parent().is_none() { host } else { dense = host.contiguous()?; &dense };
    let bytes = src.size() * std::mem::size_of::<T>();
    check(unsafe { sys::hptb_memcpy_h2d(ctx(CUDA_DEVICE)?, data.ptr().ptr as *mut c_void, src.ptr().ptr as *const c_void, bytes, stream()) })?;
    Ok(data)
}

pub fn contiguous<T, const DEVICE: usize, A>(a: &_Tensor<T, Cuda, DEVICE, A>) -> Result<_Tensor<T, Cuda, DEVICE, A>, TensorError>
where
    T: CommonBounds + HptbDtype,
    A: Allocator,
    A::Output: AllocatorOutputRetrive,
{
    let res = _Tensor::<T, Cuda, DEVICE, A>::empty(a.shape())?;
    let i = as_c(a.ptr().ptr, &a.layout());
    let mut o = as_c(res.ptr().ptr, &res.layout());
    check(unsafe { sys::hptb_copy(ctx(DEVICE)?, &i, &mut o, stream()) })?;
    Ok(res)
}

pub fn to_cpu<T, const DEVICE: usize, A>(a: &_Tensor<T, Cuda, DEVICE, A>) -> Result<_Tensor<T, Cpu, 0, <A as Allocator>::CpuAllocator>, TensorError>
where
    T: CommonBounds + HptbDtype,
    A: Allocator,
    A::Output: AllocatorOutputRetrive,
    <<A as Allocator>::CpuAllocator as Allocator>::Output: AllocatorOutputRetrive,
{
    let dense;
    let src = if a.is_contiguous() && a.parent().is_none() { a } else { dense = contiguous(a)?; &dense };
    let host = _Tensor::<T, Cpu, 0, <A as Allocator>::CpuAllocator>::empty(a.shape())?;
    let bytes = src.size() * std::mem::size_of::<T>();
    // blocking, like the reference's dtoh_sync_copy_into: the host tensor is valid on return
    check(unsafe { sys::hptb_memcpy_d2h(ctx(DEVICE)?, host.ptr().ptr as *mut c_void, src.ptr().ptr as *const c_void, bytes, stream()) })?;
    Ok(host)
}

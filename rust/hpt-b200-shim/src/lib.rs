//! hpt-b200-shim — the bodies that replace Hpt's CUDA launch templates when `hpt-cudakernels` is swapped for
//! libhpt_b200 (include/hpt_b200.h).  Each function keeps the reference's name, generic parameters and argument
//! meaning (cited per function), so the trait impls above them (`NormalBinOps`, `FloatUnaryOps`, `NormalReduce`,
//! `FloatReduce`, `IndexReduce`, `NormalizationOps`, `to_cuda::<N>()`) compile unchanged; `op_name` / `op` strings and
//! `init_val`s are still accepted and mapped to the library's enums, the `phf` kernel tables (`meta`) are ignored.
//!
//! UNVERIFIED SOURCE.  The build image has no `cargo`/`rustc`; this crate has never been compiled.  It is written
//! against Jianqoq/Hpt v0.1.3 as read in the reference checkout, and the same C entry points are exercised end to end
//! from the Python mirror (hpt_b200/tensor.py) — every host-side step below has its line-for-line twin there.
//! Meant to live at `hpt/src/backends/cuda/b200/` (it needs `pub(crate)` items of the `hpt` crate: `_Tensor`, `Cuda`).
pub mod allocator;
pub mod binary;
pub mod reduce;
pub mod softmax;
pub mod transfer;
pub mod unary;

use std::collections::HashMap;
use std::ffi::CStr;
use std::os::raw::c_void;
use std::panic::Location;
use std::sync::Mutex;

use hpt_b200_sys as sys;
use hpt_common::error::{base::TensorError, device::DeviceError, kernel::KernelError, memory::MemoryError, shape::ShapeError};
use hpt_common::layout::layout::Layout;

/// `T::HPTB_DTYPE`: the library's dtype code of a scalar type (order of hpt-types/src/dtype.rs).
pub trait HptbDtype {
    const HPTB_DTYPE: i32;
}
macro_rules! dt {
    ($($t:ty => $c:ident),*) => { $(impl HptbDtype for $t { const HPTB_DTYPE: i32 = sys::$c; })* };
}
dt!(bool => HPTB_BOOL, i8 => HPTB_I8, i16 => HPTB_I16, i32 => HPTB_I32, i64 => HPTB_I64, u8 => HPTB_U8, u16 => HPTB_U16,
    u32 => HPTB_U32, u64 => HPTB_U64, half::f16 => HPTB_F16, half::bf16 => HPTB_BF16, f32 => HPTB_F32, f64 => HPTB_F64);

/// One library context per device, created on first use (replaces the per-device PTX module cache, hpt/src/lib.rs:323-325).
pub(crate) fn ctx(device: usize) -> Result<*mut sys::hptb_ctx, TensorError> {
    static CTXS: Mutex<Option<HashMap<usize, usize>>> = Mutex::new(None);
    let mut g = CTXS.lock().unwrap();
    let map = g.get_or_insert_with(HashMap::new);
    if let Some(p) = map.get(&device) {
        return Ok(*p as *mut sys::hptb_ctx);
    }
    let mut c: *mut sys::hptb_ctx = std::ptr::null_mut();
    check(unsafe { sys::hptb_ctx_create(device as i32, &mut c) })?;
    map.insert(device, c as usize);
    Ok(c)
}

/// The stream every call is ordered on.  NULL = the legacy default stream, which is the stream cudarc 0.13's
/// `CudaDevice` launches on: observable ordering is unchanged.
pub(crate) fn stream() -> *mut c_void {
    std::ptr::null_mut()
}

/// `Layout` + device pointer → the C image (shape and strides in elements, as they are: no broadcast layout, no tables).
pub(crate) fn as_c<T: HptbDtype>(data: *mut T, layout: &Layout) -> sys::hptb_tensor {
    let nd = layout.ndim();
    let mut c = sys::hptb_tensor { data: data as *mut c_void, dtype: T::HPTB_DTYPE, ndim: nd as i32, shape: [0; 8], strides: [0; 8] };
    c.shape[..nd].copy_from_slice(layout.shape());
    c.strides[..nd].copy_from_slice(layout.strides());
    c
}

/// hptb_status → TensorError.  Shape / axis messages are already in the reference's own wording.
#[track_caller]
pub(crate) fn check(rc: sys::hptb_status) -> Result<(), TensorError> {
    if rc == sys::HPTB_OK {
        return Ok(());
    }
    let message = unsafe { CStr::from_ptr(sys::hptb_last_error()) }.to_string_lossy().into_owned();
    let location = Location::caller();
    Err(match rc {
        sys::HPTB_ERR_SHAPE | sys::HPTB_ERR_AXIS => ShapeError::BroadcastError { message, location }.into(),
        sys::HPTB_ERR_DTYPE | sys::HPTB_ERR_UNSUPPORTED | sys::HPTB_ERR_INVALID => KernelError::LaunchingError { msg: message, location }.into(),
        sys::HPTB_ERR_OOM => MemoryError::AllocationFailed { device: "cuda".to_string(), id: 0, size: 0, source: None, location }.into(),
        _ => DeviceError::CudaDriverError { message, source: None, location }.into(),
    })
}

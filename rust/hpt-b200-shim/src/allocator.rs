//! Replaces `CudaAllocator` / `_Allocator` (hpt-allocator/src/allocators/cuda.rs:41-199) and the helpers
//! `allocate_helper` / `deallocate_helper` (utils/allocate.rs:66-124, utils/deallocate.rs:11-32) for the CUDA backend.
//! The LRU keyed by the exact `Layout`, the per-allocator `HashMap<ptr, refcount>` walk on every allocation and the
//! host-synchronous `cuMemFree` go away: the library keeps size-class free lists and reuses a block in stream order.
//! Reference counting of shared storage (`insert_ptr` / `forget`) stays here — it is host bookkeeping.
use std::alloc::Layout;
use std::collections::HashMap;
use std::os::raw::c_void;
use std::sync::{Arc, Mutex};

use hpt_b200_sys as sys;
use hpt_allocator::traits::Allocator;
use hpt_common::error::base::TensorError;

use crate::{check, ctx, stream};

#[derive(Clone, Default)]
pub struct CudaAllocator {
    refs: Arc<Mutex<HashMap<(usize, usize), usize>>>, // (device, ptr) → reference count
}

impl Allocator for CudaAllocator {
    type Output = *mut u8; // the reference pairs the pointer with an Arc<CudaDevice>; the context is found by device id here
    type CpuAllocator = hpt_allocator::allocators::cpu::CpuAllocator;
    type CudaAllocator = CudaAllocator;

    fn allocate(&self, layout: Layout, device_id: usize) -> Result<Self::Output, TensorError> {
        let mut p: *mut c_void = std::ptr::null_mut();
        check(unsafe { sys::hptb_alloc(ctx(device_id)?, layout.size(), &mut p, stream()) })?;
        self.refs.lock().unwrap().insert((device_id, p as usize), 1);
        Ok(p as *mut u8)
    }

    fn allocate_zeroed(&self, layout: Layout, device_id: usize) -> Result<Self::Output, TensorError> {
        let p = self.allocate(layout, device_id)?;
        // one u8 fill over the block (set_val's replacement), ordered on the stream like everything else
        let mut t = sys::hptb_tensor { data: p as *mut c_void, dtype: sys::HPTB_U8, ndim: 1, shape: [0; 8], strides: [0; 8] };
        t.shape[0] = layout.size() as i64;
        t.strides[0] = 1;
        let zero = 0u8;
        check(unsafe { sys::hptb_fill(ctx(device_id)?, &mut t, &zero as *const u8 as *const c_void, stream()) })?;
        Ok(p)
    }

    fn deallocate(&self, ptr: *mut u8, _layout: &Layout, should_drop: bool, device_id: usize) {
        let mut refs = self.refs.lock().unwrap();
        let key = (device_id, ptr as usize);
        if let Some(n) = refs.get_mut(&key) {
            *n -= 1;
            if *n == 0 {
                refs.remove(&key);
                if should_drop {
                    if let Ok(c) = ctx(device_id) {
                        unsafe { sys::hptb_free(c, ptr as *mut c_void, stream()) };
                    }
                }
            }
        }
    }

    fn insert_ptr(&self, ptr: *mut u8, device_id: usize) {
        *self.refs.lock().unwrap().entry((device_id, ptr as usize)).or_insert(0) += 1;
    }

    fn clear(&self) {
        let devices: std::collections::HashSet<usize> = self.refs.lock().unwrap().keys().map(|k| k.0).collect();
        for d in devices {
            if let Ok(c) = ctx(d) {
                unsafe { sys::hptb_empty_cache(c) };
            }
        }
    }

    fn forget(&self, ptr: *mut u8, device_id: usize) {
        self.refs.lock().unwrap().remove(&(device_id, ptr as usize));
    }

    fn new() -> Self {
        Self::default()
    }
}

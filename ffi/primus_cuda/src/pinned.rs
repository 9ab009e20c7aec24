//! Page-locking of caller-owned host buffers.
//!
//! The `*_slices` entry points take plain host slices (the `&mut [T]` of `NttTable::transform_slice`).  Pageable memory is staged through
//! pinned bounce buffers inside the library (about half the PCIe rate); a long-lived buffer can instead be page-locked in place for the
//! lifetime of a [`Pinned`] guard, after which the same calls copy straight from / to it.
use crate::{check, sys};

/// RAII guard: the wrapped slice is registered with the CUDA driver (`pfhe_host_register`) and unregistered on drop.
/// The borrow keeps the buffer from being freed or reallocated while it is locked.
pub struct Pinned<'a, T> {
    data: &'a mut [T],
}

impl<'a, T> Pinned<'a, T> {
    pub fn new(data: &'a mut [T]) -> Self {
        let bytes = std::mem::size_of_val(data);
        check(unsafe { sys::pfhe_host_register(data.as_mut_ptr() as *mut std::ffi::c_void, bytes) }, "pfhe_host_register");
        Pinned { data }
    }
}

impl<T> std::ops::Deref for Pinned<'_, T> {
    type Target = [T];
    fn deref(&self) -> &[T] {
        self.data
    }
}

impl<T> std::ops::DerefMut for Pinned<'_, T> {
    fn deref_mut(&mut self) -> &mut [T] {
        self.data
    }
}

impl<T> Drop for Pinned<'_, T> {
    fn drop(&mut self) {
        // an error here (e.g. the context is gone at process exit) must not panic inside drop
        let _ = unsafe { sys::pfhe_host_unregister(self.data.as_mut_ptr() as *mut std::ffi::c_void) };
    }
}

/// `true` if the `*_slices` calls would stage this buffer through bounce buffers.
pub fn is_pageable<T>(data: &[T]) -> bool {
    unsafe { sys::pfhe_host_is_pageable(data.as_ptr() as *const std::ffi::c_void) != 0 }
}

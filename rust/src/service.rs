//! Streaming service mode (extension; the reference has no counterpart -- every call of src/gpu.rs:5-35,129-170 creates a
//! device, allocates its buffers and blocks in `device.poll`).  A `SigQueue` is a ring of `depth` request slots on one
//! device; each slot exposes pinned input / output slices, is submitted asynchronously and waited for individually, so
//! `depth` requests overlap on the GPU.  Thin wrapper over `sigops_queue_*` (include/sigops.h).  Not compiled in the
//! image this crate was written in (no rustc); the tested binding of the same C entry points is
//! wgpu-sigops_b200/service.py.
use crate::ffi;
use crate::ShaderFailureError;
use std::ptr;

#[derive(Clone, Copy, PartialEq, Eq, Debug)]
pub enum Curve {
    Secp256k1 = 0,
    Secp256r1 = 1,
    Ed25519 = 2,
}

pub struct SigQueue {
    q: *mut ffi::SigopsQueue,
    curve: Curve,
    max_batch: usize,
    depth: usize,
}

// the C layer serialises per-queue state with its own mutex; distinct slots may be driven from distinct threads
unsafe impl Send for SigQueue {}
unsafe impl Sync for SigQueue {}

/// The pinned arrays of one slot (`max_batch` rows each).
pub struct SlotBuffers<'a> {
    pub sigs: &'a mut [u8],
    pub msgs: &'a mut [u8],
    /// ed25519 queues only
    pub pks: Option<&'a mut [u8]>,
}

impl SigQueue {
    pub fn new(curve: Curve, device_index: i32, max_batch: usize, depth: usize) -> Result<Self, ShaderFailureError> {
        let mut q: *mut ffi::SigopsQueue = ptr::null_mut();
        let rc = unsafe { ffi::sigops_queue_create(curve as i32, device_index, max_batch, depth as i32, &mut q) };
        if rc != 0 || q.is_null() {
            return Err(ShaderFailureError);
        }
        Ok(SigQueue { q, curve, max_batch, depth })
    }

    pub fn depth(&self) -> usize {
        self.depth
    }

    fn raw(&self, slot: usize) -> Result<[*mut u8; 5], ShaderFailureError> {
        let mut p = [ptr::null_mut::<u8>(); 5];
        let rc = unsafe {
            ffi::sigops_queue_buffers(self.q, slot as i32, &mut p[0], &mut p[1], &mut p[2], &mut p[3], &mut p[4])
        };
        if rc != 0 {
            return Err(ShaderFailureError);
        }
        Ok(p)
    }

    /// Input arrays of a slot; fill the first n rows, then `submit(slot, n)`.  Must not be called while the slot is in flight.
    pub fn inputs(&mut self, slot: usize) -> Result<SlotBuffers<'_>, ShaderFailureError> {
        let p = self.raw(slot)?;
        let m = self.max_batch;
        unsafe {
            Ok(SlotBuffers {
                sigs: std::slice::from_raw_parts_mut(p[0], m * 64),
                msgs: std::slice::from_raw_parts_mut(p[1], m * 32),
                pks: if p[2].is_null() { None } else { Some(std::slice::from_raw_parts_mut(p[2], m * 32)) },
            })
        }
    }

    pub fn submit(&self, slot: usize, n: usize) -> Result<(), ShaderFailureError> {
        if unsafe { ffi::sigops_queue_submit(self.q, slot as i32, n) } != 0 {
            return Err(ShaderFailureError);
        }
        Ok(())
    }

    pub fn is_done(&self, slot: usize) -> Result<bool, ShaderFailureError> {
        let mut d = 0;
        if unsafe { ffi::sigops_queue_poll(self.q, slot as i32, &mut d) } != 0 {
            return Err(ShaderFailureError);
        }
        Ok(d != 0)
    }

    /// Blocks until the slot's request has completed; returns (keys n*64 | verdicts n, status n | empty).
    pub fn wait(&self, slot: usize) -> Result<(&[u8], &[u8]), ShaderFailureError> {
        let mut n = 0usize;
        if unsafe { ffi::sigops_queue_wait(self.q, slot as i32, &mut n, ptr::null_mut()) } != 0 {
            return Err(ShaderFailureError);
        }
        let p = self.raw(slot)?;
        let stride = if self.curve == Curve::Ed25519 { 1 } else { 64 };
        unsafe {
            let out = std::slice::from_raw_parts(p[3], n * stride);
            let st: &[u8] = if p[4].is_null() { &[] } else { std::slice::from_raw_parts(p[4], n) };
            Ok((out, st))
        }
    }
}

impl Drop for SigQueue {
    fn drop(&mut self) {
        unsafe { ffi::sigops_queue_destroy(self.q) };
    }
}

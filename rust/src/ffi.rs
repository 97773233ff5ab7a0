//! Raw bindings of include/sigops.h (the C ABI of libsigops).  One `extern "C"` item per declaration of the header.
use std::os::raw::{c_char, c_int, c_void};

extern "C" {
    pub fn sigops_init(device_ids: *const c_int, n_devices: c_int) -> c_int;
    pub fn sigops_shutdown() -> c_int;
    pub fn sigops_num_devices() -> c_int;
    pub fn sigops_last_error() -> *const c_char;
    pub fn sigops_secp256k1_ecrecover(sigs: *const u8, msgs: *const u8, n: usize, out_pubkeys: *mut u8, out_status: *mut u8) -> c_int;
    pub fn sigops_secp256r1_ecrecover(sigs: *const u8, msgs: *const u8, n: usize, out_pubkeys: *mut u8, out_status: *mut u8) -> c_int;
    pub fn sigops_ed25519_ecverify(sigs: *const u8, msgs: *const u8, pks: *const u8, n: usize, out_valid: *mut u8) -> c_int;
    pub fn sigops_ed25519_ecverify_msgs(sigs: *const u8, msg_bytes: *const u8, msg_offsets: *const u64, pks: *const u8, n: usize, flags: u32, out_valid: *mut u8) -> c_int;
    pub fn sigops_sha256_batch(data: *const u8, offsets: *const u64, n: usize, out: *mut u8) -> c_int;
    pub fn sigops_ecrecover_addresses(curve: c_int, sigs: *const u8, msg_bytes: *const u8, msg_offsets: *const u64, n: usize, out_addresses: *mut u8, out_pubkeys: *mut u8, out_status: *mut u8) -> c_int;
    pub fn sigops_precompute_bases(curve: c_int, log_limb_size: u32, out: *mut u32, inout_len: *mut usize) -> c_int;
    pub fn sigops_plan_shards(n: usize, n_devices: c_int, bounds: *mut usize, n_used: *mut c_int) -> c_int;
    pub fn sigops_last_timing(h2d_ms: *mut f64, kernel_ms: *mut f64, d2h_ms: *mut f64) -> c_int;
    pub fn sigops_kernel_launches() -> u64;
    pub fn sigops_host_alloc(bytes: usize) -> *mut c_void;
    pub fn sigops_host_free(p: *mut c_void);
    // streaming service mode (include/sigops.h, "Streaming service mode")
    pub fn sigops_queue_create(curve: c_int, device_index: c_int, max_batch: usize, depth: c_int, out: *mut *mut SigopsQueue) -> c_int;
    pub fn sigops_queue_destroy(q: *mut SigopsQueue) -> c_int;
    pub fn sigops_queue_buffers(q: *mut SigopsQueue, slot: c_int, sigs: *mut *mut u8, msgs: *mut *mut u8, pks: *mut *mut u8, out: *mut *mut u8, status: *mut *mut u8) -> c_int;
    pub fn sigops_queue_submit(q: *mut SigopsQueue, slot: c_int, n: usize) -> c_int;
    pub fn sigops_queue_submit_device(q: *mut SigopsQueue, slot: c_int, d_sigs: *const c_void, d_msgs: *const c_void, d_pks: *const c_void, n: usize, src_device: c_int, ready_event: *mut c_void) -> c_int;
    pub fn sigops_queue_slot_device(q: *mut SigopsQueue, slot: c_int) -> c_int;
    pub fn sigops_batch_on_devices(curve: c_int, device_indices: *const c_int, n_devices: c_int, sigs: *const u8, msgs: *const u8, pks: *const u8, n: usize, out: *mut u8, out_status: *mut u8) -> c_int;
    pub fn sigops_queue_poll(q: *mut SigopsQueue, slot: c_int, done: *mut c_int) -> c_int;
    pub fn sigops_queue_wait(q: *mut SigopsQueue, slot: c_int, n_done: *mut usize, device_ms: *mut f64) -> c_int;
    pub fn sigops_queue_info(q: *mut SigopsQueue, curve: *mut c_int, device_index: *mut c_int, max_batch: *mut usize, depth: *mut c_int, graph_launches: *mut u64, graph_captures: *mut u64) -> c_int;
}

/// Opaque `sigops_queue` of the C ABI.
#[repr(C)]
pub struct SigopsQueue {
    _private: [u8; 0],
}

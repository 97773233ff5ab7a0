//! `precompute` of the reference (src/precompute.rs): the 16-entry generator tables in the reference's limb format,
//! kept for API compatibility.  Computed on the CPU by libsigops; the CUDA engine never reads them.
use crate::ffi;

pub const WINDOW_SIZE: u32 = 4; // src/precompute.rs:12

fn bases(curve: i32, log_limb_size: u32) -> Vec<u32> {
    let mut len: usize = 0;
    let rc = unsafe { ffi::sigops_precompute_bases(curve, log_limb_size, std::ptr::null_mut(), &mut len) };
    assert_eq!(rc, 0, "log_limb_size must be in 11..=15");
    let mut out = vec![0u32; len];
    let rc = unsafe { ffi::sigops_precompute_bases(curve, log_limb_size, out.as_mut_ptr(), &mut len) };
    assert_eq!(rc, 0);
    out
}

/// src/precompute.rs:36-43
pub fn secp256k1_bases(log_limb_size: u32) -> Vec<u32> {
    bases(0, log_limb_size)
}

/// src/precompute.rs:45-52
pub fn secp256r1_bases(log_limb_size: u32) -> Vec<u32> {
    bases(1, log_limb_size)
}

/// src/precompute.rs:54-69
pub fn ed25519_bases(log_limb_size: u32) -> Vec<u32> {
    bases(2, log_limb_size)
}

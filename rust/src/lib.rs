//! Drop-in for `wgpu_sig_ops` (FuelLabs/wgpu-sigops): the public surface of the reference
//! (src/lib.rs, src/secp256k1_ecdsa.rs:61-66,215-219, src/secp256r1_ecdsa.rs:62-67,216-220,
//! src/ed25519_eddsa.rs:67-73,259-264, src/precompute.rs:12,36-69) over libsigops instead of wgpu + WGSL.
//! `gpu`, `shader`, `moduli`, `curve_algos`, `tests`, `benchmarks` have no counterpart: they were the wgpu runtime,
//! the WGSL templating and the crate's own test scaffolding.
pub mod ed25519_eddsa;
pub mod ffi;
pub mod precompute;
pub mod secp256k1_ecdsa;
pub mod secp256r1_ecdsa;
pub mod service;

/// This error is raised if the device path fails to execute (reference: "if the shader silently fails to execute",
/// src/lib.rs:12-14).  Here: any nonzero return code of libsigops (no CUDA device, CUDA runtime error).
#[derive(Debug, Clone)]
pub struct ShaderFailureError;

/// Reference limits and compatibility arguments: `log_limb_size` must be one the WGSL `mont_mul` supported
/// (src/wgsl/mont.wgsl:12,37) and a supplied table must have the length `precompute::*_bases` returns.
pub(crate) fn check_compat(table_limbs: Option<&Vec<u32>>, log_limb_size: u32, words_per_entry: usize) {
    assert!((11..=15).contains(&log_limb_size), "log_limb_size must be in 11..=15");
    if let Some(t) = table_limbs {
        let mut num_limbs = 256 / log_limb_size as usize;
        while num_limbs * (log_limb_size as usize) <= 256 {
            num_limbs += 1;
        }
        assert_eq!(t.len(), 16 * words_per_entry * num_limbs, "table_limbs has the wrong length");
    }
}

pub(crate) fn ecrecover_flat(
    f: unsafe extern "C" fn(*const u8, *const u8, usize, *mut u8, *mut u8) -> std::os::raw::c_int,
    sigs: &[u8],
    msgs: &[u8],
    n: usize,
) -> Result<(Vec<u8>, Vec<u8>), ShaderFailureError> {
    let mut out = vec![0u8; n * 64];
    let mut status = vec![0u8; n];
    if n == 0 {
        return Ok((out, status)); // src/secp256k1_ecdsa.rs:71-73
    }
    let rc = unsafe { f(sigs.as_ptr(), msgs.as_ptr(), n, out.as_mut_ptr(), status.as_mut_ptr()) };
    if rc != 0 {
        return Err(ShaderFailureError);
    }
    Ok((out, status))
}

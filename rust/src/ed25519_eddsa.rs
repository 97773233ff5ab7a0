//! `ed25519_eddsa` of the reference (src/ed25519_eddsa.rs): ed25519-dalek `VerifyingKey::verify` semantics.
use crate::{check_compat, ffi, ShaderFailureError};
use ed25519_dalek::{Signature, VerifyingKey};
use fuel_crypto::Message;

fn verify_flat(
    signatures: &Vec<Signature>,
    messages: &Vec<Message>,
    verifying_keys: &Vec<VerifyingKey>,
) -> Result<Vec<bool>, ShaderFailureError> {
    let n = signatures.len();
    assert_eq!(n, messages.len()); // src/ed25519_eddsa.rs:23-24
    assert_eq!(n, verifying_keys.len());
    assert!(n <= 256 * 256 * 256 * 64);
    if n == 0 {
        return Ok(vec![]); // src/ed25519_eddsa.rs:78-80
    }
    let mut s = Vec::with_capacity(n * 64);
    let mut m = Vec::with_capacity(n * 32);
    let mut k = Vec::with_capacity(n * 32);
    for i in 0..n {
        s.extend_from_slice(&signatures[i].to_bytes()); // R || s, src/ed25519_eddsa.rs:37
        k.extend_from_slice(verifying_keys[i].as_bytes()); // :38
        m.extend_from_slice(messages[i].as_slice()); // :39
    }
    let mut valid = vec![0u8; n];
    let rc = unsafe { ffi::sigops_ed25519_ecverify(s.as_ptr(), m.as_ptr(), k.as_ptr(), n, valid.as_mut_ptr()) };
    if rc != 0 {
        return Err(ShaderFailureError);
    }
    Ok(valid.iter().map(|&v| v == 1).collect()) // src/ed25519_eddsa.rs:251-254
}

/// src/ed25519_eddsa.rs:67-73
pub async fn ecverify(
    signatures: &Vec<Signature>,
    messages: &Vec<Message>,
    verifying_keys: &Vec<VerifyingKey>,
    table_limbs: &Vec<u32>,
    log_limb_size: u32,
) -> Result<Vec<bool>, ShaderFailureError> {
    check_compat(Some(table_limbs), log_limb_size, 3);
    verify_flat(signatures, messages, verifying_keys)
}

/// src/ed25519_eddsa.rs:259-264
pub async fn ecverify_single(
    signatures: &Vec<Signature>,
    messages: &Vec<Message>,
    verifying_keys: &Vec<VerifyingKey>,
    log_limb_size: u32,
) -> Result<Vec<bool>, ShaderFailureError> {
    check_compat(None, log_limb_size, 3);
    verify_flat(signatures, messages, verifying_keys)
}

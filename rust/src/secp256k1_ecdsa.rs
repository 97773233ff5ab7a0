//! `secp256k1_ecdsa` of the reference (src/secp256k1_ecdsa.rs), bodies replaced by one FFI call.
use crate::{check_compat, ecrecover_flat, ffi, ShaderFailureError};
use fuel_crypto::{Message, Signature};

fn flatten(signatures: &Vec<Signature>, messages: &Vec<Message>) -> (Vec<u8>, Vec<u8>) {
    assert_eq!(signatures.len(), messages.len()); // src/secp256k1_ecdsa.rs:21
    assert!(signatures.len() <= 256 * 256 * 256 * 64); // :22
    let mut s = Vec::with_capacity(signatures.len() * 64);
    let mut m = Vec::with_capacity(messages.len() * 32);
    for sig in signatures {
        s.extend_from_slice(sig.as_slice());
    }
    for msg in messages {
        m.extend_from_slice(msg.as_slice());
    }
    (s, m)
}

/// src/secp256k1_ecdsa.rs:61-66.  One 64-byte `X || Y` per signature; 64 zero bytes where fuel-crypto's
/// `Signature::recover` would return `Err` (the reference has no per-signature error channel).
pub async fn ecrecover(
    signatures: &Vec<Signature>,
    messages: &Vec<Message>,
    table_limbs: &Vec<u32>,
    log_limb_size: u32,
) -> Result<Vec<Vec<u8>>, ShaderFailureError> {
    check_compat(Some(table_limbs), log_limb_size, 2);
    let (keys, _) = ecrecover_with_status(signatures, messages)?;
    Ok(keys)
}

/// src/secp256k1_ecdsa.rs:215-219.  Same engine, same result (there is one fused kernel either way).
pub async fn ecrecover_single_shader(
    signatures: &Vec<Signature>,
    messages: &Vec<Message>,
    log_limb_size: u32,
) -> Result<Vec<Vec<u8>>, ShaderFailureError> {
    check_compat(None, log_limb_size, 2);
    let (keys, _) = ecrecover_with_status(signatures, messages)?;
    Ok(keys)
}

/// Extension: additionally one status byte per signature (0 = recovered, 1 = invalid signature).
pub fn ecrecover_with_status(
    signatures: &Vec<Signature>,
    messages: &Vec<Message>,
) -> Result<(Vec<Vec<u8>>, Vec<u8>), ShaderFailureError> {
    let (s, m) = flatten(signatures, messages);
    let (out, status) = ecrecover_flat(ffi::sigops_secp256k1_ecrecover, &s, &m, signatures.len())?;
    Ok((out.chunks(64).map(|c| c.to_vec()).collect(), status))
}

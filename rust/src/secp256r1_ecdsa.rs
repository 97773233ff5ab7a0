//! `secp256r1_ecdsa` of the reference (src/secp256r1_ecdsa.rs): as secp256k1 with `fuel_types::Bytes64` signatures.
use crate::{check_compat, ecrecover_flat, ffi, ShaderFailureError};
use fuel_crypto::Message;
use fuel_types::Bytes64;

/// src/secp256r1_ecdsa.rs:62-67
pub async fn ecrecover(
    signatures: &Vec<Bytes64>,
    messages: &Vec<Message>,
    table_limbs: &Vec<u32>,
    log_limb_size: u32,
) -> Result<Vec<Vec<u8>>, ShaderFailureError> {
    check_compat(Some(table_limbs), log_limb_size, 2);
    let (keys, _) = ecrecover_with_status(signatures, messages)?;
    Ok(keys)
}

/// src/secp256r1_ecdsa.rs:216-220
pub async fn ecrecover_single_shader(
    signatures: &Vec<Bytes64>,
    messages: &Vec<Message>,
    log_limb_size: u32,
) -> Result<Vec<Vec<u8>>, ShaderFailureError> {
    check_compat(None, log_limb_size, 2);
    let (keys, _) = ecrecover_with_status(signatures, messages)?;
    Ok(keys)
}

pub fn ecrecover_with_status(
    signatures: &Vec<Bytes64>,
    messages: &Vec<Message>,
) -> Result<(Vec<Vec<u8>>, Vec<u8>), ShaderFailureError> {
    assert_eq!(signatures.len(), messages.len()); // src/secp256r1_ecdsa.rs:22
    assert!(signatures.len() <= 256 * 256 * 256 * 64);
    let mut s = Vec::with_capacity(signatures.len() * 64);
    let mut m = Vec::with_capacity(messages.len() * 32);
    for sig in signatures {
        s.extend_from_slice(sig.as_slice());
    }
    for msg in messages {
        m.extend_from_slice(msg.as_slice());
    }
    let (out, status) = ecrecover_flat(ffi::sigops_secp256r1_ecrecover, &s, &m, signatures.len())?;
    Ok((out.chunks(64).map(|c| c.to_vec()).collect(), status))
}

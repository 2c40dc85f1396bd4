//! Batched witness hints for plonky2x backed by libbsx (B200 kernels).
//!
//! SOURCE FOR THE REFERENCE SIDE -- not built in this repository (its image has no Rust toolchain).  Drop this file into
//! `plonky2x/core/src/frontend/hint/` next to the existing hints, add `bsx-sys` (bindings/rust/bsx-sys) as a dependency and
//! register the three types in `HintRegistry` (`backend/circuit/serialization/hints.rs:401-423`).
//!
//! What changes in plonky2x: `curta_constrain_hashes` (`frontend/hash/curta/builder.rs:26-49`) and
//! `curta_constrain_ec_ops` (`frontend/ecc/curve25519/curta/builder.rs:18-64`) loop over the queued requests and register
//! ONE hint per request (`HashDigestHint`, `EcOpResultHint`), each running scalar CPU code.  With the hints below the loop
//! writes every request into one input stream and registers ONE hint per accelerator; its `hint()` packs the requests,
//! calls libbsx once and writes the responses back in request order.  Nothing in the circuits' operator surface
//! (`curta_sha256*`, `curta_sha512*`, `curta_25519_*`, `curta_eddsa_verify_sigs*`, the Tendermint Merkle methods,
//! `circuits/builder.rs`, `vars.rs`) changes: those methods only enqueue requests.
//!
//! Byte conventions are the ones `include/bsx.h` documents and `tests/` check bit for bit against the reference's
//! fixtures: SHA digests are the big-endian state words (`sha256/curta.rs:81-92`; SHA-512 words as (low, high) u32 limbs,
//! `sha512/curta.rs:68-88`), field elements are 32 little-endian bytes = the 16 u16 limbs of `FieldVariable`.

use std::cell::RefCell;

use bsx_sys::*;
use serde::{Deserialize, Serialize};

use crate::frontend::curta::ec::point::{AffinePointVariable, CompressedEdwardsYVariable};
use crate::frontend::curta::field::variable::FieldVariable;
use crate::frontend::ecc::curve25519::curta::Curve;
use crate::frontend::hint::simple::hint::Hint;
use crate::prelude::*;

thread_local! {
    /// One ctx per witness thread: a `bsx_ctx` is not thread-safe, exactly like the reference's single-threaded
    /// generator loop (`backend/circuit/witness.rs:144-199`); rayon callers get one ctx each.
    static CTX: RefCell<*mut bsx_ctx> = RefCell::new(std::ptr::null_mut());
}

pub(crate) fn with_ctx<T>(f: impl FnOnce(*mut bsx_ctx) -> T) -> T {
    CTX.with(|c| {
        let mut c = c.borrow_mut();
        if c.is_null() {
            let device = std::env::var("BSX_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
            let rc = unsafe { bsx_init(device, &mut *c) };
            assert_eq!(rc, BSX_OK, "bsx_init failed ({rc}): no CUDA device -- libbsx has no CPU fallback");
        }
        f(*c)
    })
}

pub(crate) fn check(ctx: *mut bsx_ctx, rc: i32, what: &str) {
    if rc != BSX_OK {
        let msg = unsafe { std::ffi::CStr::from_ptr(bsx_last_error(ctx)) }.to_string_lossy().into_owned();
        panic!("{what} failed ({rc}): {msg}"); // hints report failure by panicking (simple/generator.rs:74-79)
    }
}

/// How one queued hash request is laid out in the input stream (recorded at build time).
#[derive(Clone, Debug, Serialize, Deserialize)]
pub enum BatchedHashRequest {
    /// `HashRequest::Fixed(msg)`: `len` bytes follow.
    Fixed(usize),
    /// `HashRequest::Variable(msg, len, _)`: a length `Variable`, then a buffer of `buf_len` bytes; only the first
    /// `len` are hashed (`digest_hint.rs:31-33`).
    Variable(usize),
}

/// All SHA-256 requests of one circuit in one call: replaces the per-request `HashDigestHint<SHA256, ..>`.
#[derive(Clone, Debug, Serialize, Deserialize)]
pub struct BatchedSha256DigestHint {
    pub requests: Vec<BatchedHashRequest>,
}

fn read_requests<L: PlonkParameters<D>, const D: usize>(
    requests: &[BatchedHashRequest],
    input_stream: &mut ValueStream<L, D>,
) -> (Vec<u8>, Vec<u32>) {
    let (mut msgs, mut offsets) = (Vec::<u8>::new(), vec![0u32]);
    for r in requests {
        match r {
            BatchedHashRequest::Fixed(len) => msgs.extend(input_stream.read_vec::<ByteVariable>(*len)),
            BatchedHashRequest::Variable(buf_len) => {
                let len = input_stream.read_value::<Variable>().as_canonical_u64() as usize;
                let buf = input_stream.read_vec::<ByteVariable>(*buf_len);
                msgs.extend_from_slice(&buf[..len]);
            }
        }
        offsets.push(msgs.len() as u32);
    }
    (msgs, offsets)
}

impl<L: PlonkParameters<D>, const D: usize> Hint<L, D> for BatchedSha256DigestHint {
    fn hint(&self, input_stream: &mut ValueStream<L, D>, output_stream: &mut ValueStream<L, D>) {
        let (msgs, offsets) = read_requests(&self.requests, input_stream);
        let n = self.requests.len();
        let mut digests = vec![0u8; 32 * n];
        with_ctx(|ctx| {
            let rc = unsafe { bsx_sha256_batch(ctx, msgs.as_ptr(), offsets.as_ptr(), n as u32, digests.as_mut_ptr()) };
            check(ctx, rc, "bsx_sha256_batch");
        });
        for d in digests.chunks_exact(32) {
            // [U32Variable; 8]: big-endian state words
            let words: [u32; 8] = core::array::from_fn(|i| u32::from_be_bytes(d[4 * i..4 * i + 4].try_into().unwrap()));
            output_stream.write_value::<[U32Variable; 8]>(words);
        }
    }
}

/// All SHA-512 requests of one circuit: replaces `HashDigestHint<SHA512, ..>`.
#[derive(Clone, Debug, Serialize, Deserialize)]
pub struct BatchedSha512DigestHint {
    pub requests: Vec<BatchedHashRequest>,
}

impl<L: PlonkParameters<D>, const D: usize> Hint<L, D> for BatchedSha512DigestHint {
    fn hint(&self, input_stream: &mut ValueStream<L, D>, output_stream: &mut ValueStream<L, D>) {
        let (msgs, offsets) = read_requests(&self.requests, input_stream);
        let n = self.requests.len();
        let mut digests = vec![0u8; 64 * n];
        with_ctx(|ctx| {
            let rc = unsafe { bsx_sha512_batch(ctx, msgs.as_ptr(), offsets.as_ptr(), n as u32, digests.as_mut_ptr()) };
            check(ctx, rc, "bsx_sha512_batch");
        });
        for d in digests.chunks_exact(64) {
            // [U64Variable; 8]: big-endian state words, each written as a u64 (two u32 limbs, low first)
            let words: [u64; 8] = core::array::from_fn(|i| u64::from_be_bytes(d[8 * i..8 * i + 8].try_into().unwrap()));
            output_stream.write_value::<[U64Variable; 8]>(words);
        }
    }
}

/// Every EC request `curta_eddsa_verify_sigs[_conditional]` issues for `num_sigs` signatures
/// (`ecc/curve25519/ed25519/eddsa.rs:161-203`), in its order: per signature ScalarMul(s, G), Decompress(A), IsValid,
/// ScalarMul(h, A), Decompress(R), IsValid, Add(R, hA) -- plus the `h = digest mod l` quotient/remainder that
/// `BigUintDivRemGenerator` (`uint/num/biguint/mod.rs:451-488`) produces.  Replaces 7 `EcOpResultHint`s, one
/// `HashDigestHint<SHA512>` and one div/rem generator per signature.
#[derive(Clone, Debug, Serialize, Deserialize)]
pub struct BatchedEd25519Hint {
    pub num_sigs: usize,
    /// bytes of every message buffer (124 in tendermintx)
    pub msg_stride: usize,
}

impl<L: PlonkParameters<D>, const D: usize> Hint<L, D> for BatchedEd25519Hint {
    fn hint(&self, input_stream: &mut ValueStream<L, D>, output_stream: &mut ValueStream<L, D>) {
        let n = self.num_sigs;
        let (mut pks, mut sigs, mut msgs, mut lens) = (vec![0u8; 32 * n], vec![0u8; 64 * n], vec![0u8; self.msg_stride * n], vec![0u32; n]);
        for i in 0..n {
            // the inputs of one lane, as the gadget holds them: compressed pubkey, R, s (U256), message, length
            let pk = input_stream.read_value::<CompressedEdwardsYVariable>();
            let r = input_stream.read_value::<CompressedEdwardsYVariable>();
            let s = input_stream.read_value::<U256Variable>();
            let msg = input_stream.read_vec::<ByteVariable>(self.msg_stride);
            lens[i] = input_stream.read_value::<U32Variable>();
            pks[32 * i..32 * i + 32].copy_from_slice(pk.as_bytes());
            sigs[64 * i..64 * i + 32].copy_from_slice(r.as_bytes());
            s.to_little_endian(&mut sigs[64 * i + 32..64 * i + 64]);
            msgs[self.msg_stride * i..self.msg_stride * (i + 1)].copy_from_slice(&msg);
        }
        let mut out = vec![0u8; BSX_SIG_OUT_BYTES as usize * n];
        with_ctx(|ctx| {
            let rc = unsafe {
                bsx_ed25519_batch(ctx, n as u32, pks.as_ptr(), sigs.as_ptr(), msgs.as_ptr(), self.msg_stride as u32, lens.as_ptr(),
                                  std::ptr::null(), out.as_mut_ptr())
            };
            check(ctx, rc, "bsx_ed25519_batch");
        });
        type Fe = FieldVariable<<Curve as starkyx::chip::ec::EllipticCurveParameters>::BaseField>;
        let fe = |b: &[u8]| num_bigint::BigUint::from_bytes_le(b);
        for rec in out.chunks_exact(BSX_SIG_OUT_BYTES as usize) {
            let flags = rec[520] as u32;
            // the reference's `decompress` panics on a point that is not on the curve
            assert!(flags & BSX_SIG_A_OK != 0 && flags & BSX_SIG_R_OK != 0, "Ed25519 point does not decompress");
            let point = |o: usize| (fe(&rec[o..o + 32]), fe(&rec[o + 32..o + 64]));
            output_stream.write_value::<[U64Variable; 8]>(core::array::from_fn(|i| u64::from_be_bytes(rec[8 * i..8 * i + 8].try_into().unwrap()))); // SHA-512 digest
            output_stream.write_value::<U512Variable>(ethers::types::U512::from_little_endian(&[&rec[96..136], &[0u8; 24][..]].concat()));   // div
            output_stream.write_value::<U512Variable>(ethers::types::U512::from_little_endian(&[&rec[64..96], &[0u8; 32][..]].concat()));    // rem = h
            output_stream.write_value::<AffinePointVariable<Curve>>(point(136).into());   // s G
            output_stream.write_value::<AffinePointVariable<Curve>>(point(200).into());   // A
            output_stream.write_value::<Fe>(fe(&rec[264..296]));                          // root of A
            output_stream.write_value::<AffinePointVariable<Curve>>(point(296).into());   // h A
            output_stream.write_value::<AffinePointVariable<Curve>>(point(360).into());   // R
            output_stream.write_value::<Fe>(fe(&rec[424..456]));                          // root of R
            output_stream.write_value::<AffinePointVariable<Curve>>(point(456).into());   // R + h A
        }
    }
}

//! Device-side input shaping for the Blobstream X input fetcher, backed by libbsx (B200 kernels).
//!
//! SOURCE FOR THE REFERENCE SIDE -- not built in this repository (its image has no Rust toolchain).  It belongs next to
//! `circuits/input.rs` of blobstreamx: `DataCommitmentInputs::get_data_commitment_inputs` (`circuits/input.rs:149-271`)
//! encodes 14 protobuf fields and hashes 27 SHA-256 messages per header on the host to build the two inclusion proofs of
//! every header of a map job; with the functions below the fetcher only decodes the RPC's JSON into `bsx_header_fields`
//! and gets back the byte arrays of `DataCommitmentProofValueType` for all 32 jobs of a range -- or leaves them on the
//! device for `bsx_header_range_dev`.  The skip / step side (`tendermintx/circuits/input/conversion.rs:20-240`) is
//! `shape_validators` below.  Parity of every array with the host code is what `tests/test_gpu_inputs.py` and
//! `tests/test_gpu_encode.py` check against the reference's fixtures.

use bsx_sys::*;
use tendermint::block::{signed_header::SignedHeader, CommitSig, Header};
use tendermint::validator::Info;

use super::batched_hints::{check, with_ctx}; // the thread-local ctx of the hints

/// `Header` -> the fixed layout `bsx_encode_headers` takes (`include/bsx.h`).  Only copies: the encoders run on the device.
pub fn header_fields(h: &Header) -> bsx_header_fields {
    let mut f: bsx_header_fields = unsafe { std::mem::zeroed() };
    f.version_block = h.version.block;
    f.version_app = h.version.app;
    f.height = h.height.value();
    let t: tendermint_proto::google::protobuf::Timestamp = h.time.into();
    f.time_seconds = t.seconds;
    f.time_nanos = t.nanos as u32;
    let id = h.chain_id.as_bytes();
    f.chain_id_len = id.len() as u32;
    f.chain_id[..id.len()].copy_from_slice(id);
    if let Some(b) = h.last_block_id {
        f.has_last_block_id = 1;
        f.parts_total = b.part_set_header.total;
        f.last_block_hash.copy_from_slice(b.hash.as_bytes());
        f.parts_hash.copy_from_slice(b.part_set_header.hash.as_bytes());
    }
    // last_commit_hash, data_hash, validators_hash, next_validators_hash, consensus_hash, app_hash,
    // last_results_hash, evidence_hash, proposer_address -- `unwrap_or_default()` as in generate_proofs_from_header
    let hashes: [Vec<u8>; 9] = [
        h.last_commit_hash.map(|x| x.as_bytes().to_vec()).unwrap_or_default(),
        h.data_hash.map(|x| x.as_bytes().to_vec()).unwrap_or_default(),
        h.validators_hash.as_bytes().to_vec(),
        h.next_validators_hash.as_bytes().to_vec(),
        h.consensus_hash.as_bytes().to_vec(),
        h.app_hash.as_bytes().to_vec(),
        h.last_results_hash.map(|x| x.as_bytes().to_vec()).unwrap_or_default(),
        h.evidence_hash.map(|x| x.as_bytes().to_vec()).unwrap_or_default(),
        h.proposer_address.as_bytes().to_vec(),
    ];
    for (k, v) in hashes.iter().enumerate() {
        f.hash_len[k] = v.len() as u8;
        f.hashes[k][..v.len()].copy_from_slice(v);
    }
    f
}

/// The arrays of `DataCommitmentProofValueType` for the `n_jobs` map jobs of ONE range, byte for byte what
/// `get_data_commitment_inputs` returns job by job: each job is clamped to `latest_safe` (= latest_block - 2), not to the
/// range's `end`; slots beyond `latest_safe` are zero.
pub struct RangeMapInputs {
    pub dh_leaf: Vec<u8>,       // n_jobs * B * 34
    pub dh_aunts: Vec<u8>,      // n_jobs * B * 4 * 32
    pub lb_leaf: Vec<u8>,       // n_jobs * B * 72
    pub lb_aunts: Vec<u8>,      // n_jobs * B * 4 * 32
    pub start_headers: Vec<u8>, // n_jobs * 32
    pub end_headers: Vec<u8>,   // n_jobs * 32
    pub start_header: [u8; 32],
    pub end_header: [u8; 32],
}

/// `headers`: blocks `start ..= start + n_jobs * B` (blocks beyond `latest_safe`: any value, they are ignored).
pub fn shape_range(headers: &[Header], start: u64, end: u64, latest_safe: u64, n_jobs: u32, b: u32) -> RangeMapInputs {
    let n = headers.len();
    assert_eq!(n as u32, n_jobs * b + 1);
    let fields: Vec<bsx_header_fields> = headers.iter().map(header_fields).collect();
    let mut records = vec![0u8; n * BSX_HEADER_LEAVES_BYTES as usize];
    let slots = (n_jobs * b) as usize;
    let mut out = RangeMapInputs {
        dh_leaf: vec![0; slots * 34],
        dh_aunts: vec![0; slots * 128],
        lb_leaf: vec![0; slots * 72],
        lb_aunts: vec![0; slots * 128],
        start_headers: vec![0; n_jobs as usize * 32],
        end_headers: vec![0; n_jobs as usize * 32],
        start_header: [0; 32],
        end_header: [0; 32],
    };
    let mut fail = 0u32;
    with_ctx(|ctx| unsafe {
        check(ctx, bsx_encode_headers(ctx, n as u32, fields.as_ptr(), records.as_mut_ptr()), "bsx_encode_headers");
        check(
            ctx,
            bsx_header_range_inputs(
                ctx, 1, n_jobs, b, records.as_ptr(), &start, &end, &latest_safe, out.dh_leaf.as_mut_ptr(), out.dh_aunts.as_mut_ptr(),
                out.lb_leaf.as_mut_ptr(), out.lb_aunts.as_mut_ptr(), out.start_headers.as_mut_ptr(),
                out.end_headers.as_mut_ptr(), out.start_header.as_mut_ptr(), out.end_header.as_mut_ptr(), &mut fail,
            ),
            "bsx_header_range_inputs",
        );
    });
    // the host code panics in `get_inclusion_proof` when a proven field does not have the circuit's fixed size
    assert_eq!(fail & BSX_FAIL_INPUT_LEAF, 0, "data_hash / last_block_id field is not 34 / 72 bytes");
    out
}

fn sig_slot(v: &Info, s: &CommitSig) -> bsx_commit_sig_in {
    let mut r: bsx_commit_sig_in = unsafe { std::mem::zeroed() };
    r.pubkey.copy_from_slice(&v.pub_key.to_bytes());
    r.voting_power = v.power();
    r.address.copy_from_slice(v.address.as_bytes());
    match s {
        CommitSig::BlockIdFlagAbsent => r.block_id_flag = 1,
        CommitSig::BlockIdFlagCommit { validator_address, timestamp, signature }
        | CommitSig::BlockIdFlagNil { validator_address, timestamp, signature } => {
            r.block_id_flag = if s.is_commit() { 2 } else { 3 };
            r.sig_address.copy_from_slice(validator_address.as_bytes());
            let t: tendermint_proto::google::protobuf::Timestamp = (*timestamp).into();
            r.ts_seconds = t.seconds;
            r.ts_nanos = t.nanos as u32;
            if let Some(sig) = signature {
                r.signature.copy_from_slice(sig.as_bytes());
            }
        }
    }
    r
}

/// `get_validator_data_from_block` + `update_present_on_trusted_header` (+ `validator_hash_field_from_block` for the
/// trusted set) of verify_skip: -> (ValidatorVariable records N * 240 bytes, trusted pubkeys / powers / byte lengths).
pub fn shape_validators<const N: usize>(
    target: &SignedHeader, target_validators: &[Info], trusted_validators: &[Info],
) -> (Vec<u8>, Vec<u8>, Vec<u64>, Vec<u32>) {
    let c = &target.commit;
    let mut cm: bsx_commit_in = unsafe { std::mem::zeroed() };
    cm.height = c.height.value();
    cm.round = c.round.value();
    cm.n_signatures = c.signatures.len() as u32;
    cm.has_block_id = 1;
    cm.block_hash.copy_from_slice(c.block_id.hash.as_bytes());
    cm.parts_hash.copy_from_slice(c.block_id.part_set_header.hash.as_bytes());
    cm.parts_total = c.block_id.part_set_header.total;
    let id = target.header.chain_id.as_bytes();
    cm.chain_id_len = id.len() as u32;
    cm.chain_id[..id.len()].copy_from_slice(id);
    let zero: bsx_commit_sig_in = unsafe { std::mem::zeroed() };
    let mut tg = vec![zero; N];
    for (i, (v, s)) in target_validators.iter().zip(c.signatures.iter()).enumerate() {
        tg[i] = sig_slot(v, s);
    }
    let mut tr = vec![zero; N];
    for (i, v) in trusted_validators.iter().enumerate() {
        tr[i].pubkey.copy_from_slice(&v.pub_key.to_bytes());
        tr[i].voting_power = v.power();
        tr[i].address.copy_from_slice(v.address.as_bytes());
    }
    let mut tcm: bsx_commit_in = unsafe { std::mem::zeroed() };
    tcm.n_signatures = trusted_validators.len() as u32;
    let (mut records, mut pks, mut powers, mut lens) = (vec![0u8; N * 240], vec![0u8; N * 32], vec![0u64; N], vec![0u32; N]);
    let (mut fail, n_target, n_trusted) = (0u32, target_validators.len() as u32, trusted_validators.len() as u32);
    with_ctx(|ctx| unsafe {
        let null8 = std::ptr::null_mut::<u8>();
        check(
            ctx,
            bsx_validator_records(ctx, 1, N as u32, &cm, tg.as_ptr(), records.as_mut_ptr(), null8, std::ptr::null_mut(), std::ptr::null_mut(), &mut fail),
            "bsx_validator_records",
        );
        assert_eq!(fail, 0, "sign bytes longer than VALIDATOR_MESSAGE_BYTES_LENGTH_MAX"); // `try_into().unwrap()` on the host
        check(
            ctx,
            bsx_validator_records(ctx, 1, N as u32, &tcm, tr.as_ptr(), null8, pks.as_mut_ptr(), powers.as_mut_ptr(), lens.as_mut_ptr(), &mut fail),
            "bsx_validator_records (trusted hash fields)",
        );
        check(
            ctx,
            bsx_present_on_trusted(ctx, 1, N as u32, tg.as_ptr(), &n_target, tr.as_ptr(), &n_trusted, records.as_mut_ptr(), &mut fail),
            "bsx_present_on_trusted",
        );
        assert_eq!(fail & BSX_FAIL_INPUT_THRESHOLD, 0, "shared voting power is less than threshold"); // the host asserts
    });
    (records, pks, powers, lens)
}

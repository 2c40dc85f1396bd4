"""Host-side input shaping: chain data -> the flat arrays the hot path consumes.

Mirrors the reference's off-chain input code (which is network/JSON glue and stays on the host):
  * DataCommitmentInputFetcher.get_data_commitment_inputs   BX/circuits/input.rs:149-271
  * InputDataFetcher.get_step_inputs / get_skip_inputs      TX/input/mod.rs:317-530
  * get_validator_data_from_block / validator_hash_field_from_block / update_present_on_trusted_header
                                                            TX/input/conversion.rs:59-240
  * generate_proofs_from_header                             TX/input/tendermint_utils.rs:374-393
Header hashing here uses hashlib (input preparation, not the measured path); everything the
circuits themselves hash goes through the CUDA library.
"""
from __future__ import annotations

import base64
import calendar
import hashlib
import re
import time
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# sizes (TX/consts.rs:4-37, BX/circuits/consts.rs:4-23)
VALIDATOR_SET_SIZE_MAX = 100
VALIDATOR_MESSAGE_BYTES_LENGTH_MAX = 124
VALIDATOR_BYTE_LENGTH_MAX = 46
PROTOBUF_CHAIN_ID_SIZE_BYTES = 52
PROTOBUF_HASH_SIZE_BYTES = 34
PROTOBUF_BLOCK_ID_SIZE_BYTES = 72
HEADER_PROOF_DEPTH = 4
CHAIN_ID_INDEX, BLOCK_HEIGHT_INDEX, LAST_BLOCK_ID_INDEX = 1, 2, 4
DATA_HASH_INDEX, VALIDATORS_HASH_INDEX, NEXT_VALIDATORS_HASH_INDEX = 6, 7, 8

VAL_IN_BYTES = 240  # validator record, see include/bsx.h
HASHPROOF_BYTES = 34 + 128
BLOCKIDPROOF_BYTES = 72 + 128

# PX/frontend/ecc/curve25519/ed25519/eddsa.rs:27-42
DUMMY_PUBLIC_KEY = bytes([138, 136, 227, 221, 116, 9, 241, 149, 253, 82, 219, 45, 60, 186, 93, 114, 202, 103, 9, 191,
                          29, 148, 18, 27, 243, 116, 136, 1, 180, 15, 111, 92])
DUMMY_SIGNATURE = bytes([55, 20, 104, 158, 84, 120, 194, 17, 6, 237, 157, 164, 85, 88, 158, 137, 187, 119, 187, 240,
                         159, 73, 80, 63, 133, 162, 74, 91, 48, 53, 6, 138, 1, 41, 22, 121, 249, 46, 198, 145, 155, 102,
                         3, 210, 168, 135, 173, 55, 252, 72, 45, 126, 169, 178, 191, 7, 153, 67, 112, 90, 150, 33, 140, 7])


def _uvarint(n: int) -> bytes:
    b = bytearray()
    while n >= 0x80:
        b.append((n & 0x7F) | 0x80)
        n >>= 7
    b.append(n)
    return bytes(b)


def _ld(tag: int, payload: bytes) -> bytes:  # length-delimited field, omitted when empty
    return (bytes([tag]) + _uvarint(len(payload)) + payload) if payload else b""


def _vi(tag: int, n: int) -> bytes:  # varint field, omitted when zero
    return (bytes([tag]) + _uvarint(n)) if n else b""


_TS = re.compile(r"^(\d+)-(\d+)-(\d+)T(\d+):(\d+):(\d+)(?:\.(\d+))?Z$")


def _timestamp(ts) -> bytes:
    if isinstance(ts, tuple):
        secs, nanos = ts
    else:
        m = _TS.match(ts)
        if not m:
            raise ValueError(f"bad RFC3339 time {ts!r}")
        y, mo, d, hh, mm, ss = (int(x) for x in m.groups()[:6])
        secs = calendar.timegm((y, mo, d, hh, mm, ss))
        nanos = int((m.group(7) or "").ljust(9, "0")[:9] or 0)
    return _vi(0x08, secs) + _vi(0x10, nanos)


def _block_id(bid: Optional[dict]) -> bytes:
    if not bid or not bid.get("hash"):
        return b""
    parts = _vi(0x08, int(bid["parts"]["total"])) + _ld(0x12, bytes.fromhex(bid["parts"]["hash"]))
    return _ld(0x0A, bytes.fromhex(bid["hash"])) + _ld(0x12, parts)


def header_leaves(h: dict) -> List[bytes]:
    """Protobuf encodings of the 14 header fields (TX/input/tendermint_utils.rs:374-393)."""
    hx = lambda k: bytes.fromhex(h.get(k) or "")
    return [
        _vi(0x08, int(h["version"]["block"])) + _vi(0x10, int(h["version"].get("app") or 0)),
        _ld(0x0A, h["chain_id"].encode()),
        _vi(0x08, int(h["height"])),
        _timestamp(h["time"]),
        _block_id(h.get("last_block_id")),
        _ld(0x0A, hx("last_commit_hash")),
        _ld(0x0A, hx("data_hash")),
        _ld(0x0A, hx("validators_hash")),
        _ld(0x0A, hx("next_validators_hash")),
        _ld(0x0A, hx("consensus_hash")),
        _ld(0x0A, hx("app_hash")),
        _ld(0x0A, hx("last_results_hash")),
        _ld(0x0A, hx("evidence_hash")),
        _ld(0x0A, hx("proposer_address")),
    ]


def _h(b: bytes) -> bytes:
    return hashlib.sha256(b).digest()


@dataclass
class HeaderTree:
    """The 14-leaf Tendermint tree of one header: 8|6 -> (4|4) | (4|2)."""
    leaves: List[bytes]
    root: bytes
    _levels: Tuple

    @staticmethod
    def build(leaves: Sequence[bytes]) -> "HeaderTree":
        assert len(leaves) == 14
        l0 = [_h(b"\x00" + x) for x in leaves]
        l1 = [_h(b"\x01" + l0[i] + l0[i + 1]) for i in range(0, 14, 2)]        # 7 nodes
        l2 = [_h(b"\x01" + l1[0] + l1[1]), _h(b"\x01" + l1[2] + l1[3]), _h(b"\x01" + l1[4] + l1[5]), l1[6]]
        l3 = [_h(b"\x01" + l2[0] + l2[1]), _h(b"\x01" + l2[2] + l2[3])]
        root = _h(b"\x01" + l3[0] + l3[1])
        return HeaderTree(list(leaves), root, (l0, l1, l2, l3))

    def aunts(self, index: int) -> List[bytes]:
        """Depth-4 proof for leaves 0..11 (leaves 12,13 sit at depth 3 and are never proven)."""
        assert 0 <= index < 12
        l0, l1, l2, l3 = self._levels
        return [l0[index ^ 1], l1[(index >> 1) ^ 1], l2[(index >> 2) ^ 1], l3[(index >> 3) ^ 1]]


HEADER_LEAVES_BYTES = 512   # include/bsx.h BSX_HEADER_LEAVES_BYTES


def pack_header_record(leaves: Sequence[bytes]) -> np.ndarray:
    """The 14 encoded header fields as the record bsx_header_trees / bsx_header_range_inputs take:
    bytes [0,14) = field lengths, bytes [16, 16+sum) = the fields back to back."""
    assert len(leaves) == 14 and all(len(x) < 256 for x in leaves)
    body = b"".join(leaves)
    if 16 + len(body) > HEADER_LEAVES_BYTES:
        raise ValueError(f"encoded header fields are {len(body)} bytes, the record holds {HEADER_LEAVES_BYTES - 16}")
    rec = np.zeros(HEADER_LEAVES_BYTES, np.uint8)
    rec[:14] = [len(x) for x in leaves]
    rec[16:16 + len(body)] = np.frombuffer(body, np.uint8)
    return rec


def pack_range_headers(trees: Dict[int, "HeaderTree"], start: int, n_jobs: int, batch_size: int) -> np.ndarray:
    """Records of blocks start .. start + n_jobs*batch_size (missing blocks beyond the chain tip stay zero)."""
    out = np.zeros((n_jobs * batch_size + 1, HEADER_LEAVES_BYTES), np.uint8)
    for o in range(n_jobs * batch_size + 1):
        t = trees.get(start + o)
        if t is not None:
            out[o] = pack_header_record(t.leaves)
    return out


# fixed-layout inputs of the device-side encoders (include/bsx.h: bsx_header_fields, bsx_commit_in, bsx_commit_sig_in)
HEADER_FIELDS_DTYPE = np.dtype([
    ("version_block", "<u8"), ("version_app", "<u8"), ("height", "<u8"), ("time_seconds", "<i8"), ("time_nanos", "<u4"),
    ("chain_id_len", "<u4"), ("chain_id", "u1", 56), ("parts_total", "<u4"), ("has_last_block_id", "u1"), ("hash_len", "u1", 9),
    ("_pad", "u1", 2), ("last_block_hash", "u1", 32), ("parts_hash", "u1", 32), ("hashes", "u1", (9, 32))])
COMMIT_DTYPE = np.dtype([
    ("height", "<u8"), ("round", "<u4"), ("n_signatures", "<u4"), ("block_hash", "u1", 32), ("parts_hash", "u1", 32),
    ("parts_total", "<u4"), ("chain_id_len", "<u4"), ("chain_id", "u1", 56), ("has_block_id", "u1"), ("_pad", "u1", 7)])
COMMIT_SIG_DTYPE = np.dtype([
    ("pubkey", "u1", 32), ("signature", "u1", 64), ("voting_power", "<u8"), ("ts_seconds", "<i8"), ("ts_nanos", "<u4"),
    ("block_id_flag", "u1"), ("_pad", "u1", 3), ("address", "u1", 20), ("sig_address", "u1", 20)])
assert HEADER_FIELDS_DTYPE.itemsize == 464 and COMMIT_DTYPE.itemsize == 152 and COMMIT_SIG_DTYPE.itemsize == 160

_HEADER_HASH_KEYS = ("last_commit_hash", "data_hash", "validators_hash", "next_validators_hash", "consensus_hash", "app_hash",
                     "last_results_hash", "evidence_hash", "proposer_address")


def parse_time(ts) -> Tuple[int, int]:
    """RFC 3339 (as in the RPC's JSON) or (seconds, nanos) -> (seconds, nanos)."""
    if isinstance(ts, tuple):
        return int(ts[0]), int(ts[1])
    m = _TS.match(ts)
    if not m:
        raise ValueError(f"bad RFC3339 time {ts!r}")
    y, mo, d, hh, mm, ss = (int(x) for x in m.groups()[:6])
    return calendar.timegm((y, mo, d, hh, mm, ss)), int((m.group(7) or "").ljust(9, "0")[:9] or 0)


def _put(dst: np.ndarray, b: bytes) -> int:
    dst[: len(b)] = np.frombuffer(b, np.uint8)
    return len(b)


def pack_header_fields(headers: Sequence[dict]) -> np.ndarray:
    """Decoded headers (the RPC's JSON dicts) -> bsx_header_fields records for `encode_headers`."""
    out = np.zeros(len(headers), HEADER_FIELDS_DTYPE)
    for r, h in zip(out, headers):
        r["version_block"], r["version_app"] = int(h["version"]["block"]), int(h["version"].get("app") or 0)
        r["height"] = int(h["height"])
        r["time_seconds"], r["time_nanos"] = parse_time(h["time"])
        r["chain_id_len"] = _put(r["chain_id"], h["chain_id"].encode())
        bid = h.get("last_block_id")
        if bid and bid.get("hash"):
            r["has_last_block_id"], r["parts_total"] = 1, int(bid["parts"]["total"])
            _put(r["last_block_hash"], bytes.fromhex(bid["hash"]))
            _put(r["parts_hash"], bytes.fromhex(bid["parts"]["hash"]))
        for k, key in enumerate(_HEADER_HASH_KEYS):
            r["hash_len"][k] = _put(r["hashes"][k], bytes.fromhex(h.get(key) or ""))
    return out


def _b64(x) -> bytes:
    return bytes(x) if isinstance(x, (bytes, bytearray)) else base64.b64decode(x)


def pack_commit(header: dict, commit: dict, validators: Sequence[dict], n_max: int = VALIDATOR_SET_SIZE_MAX):
    """One commit and its validator set -> (bsx_commit_in record, n_max bsx_commit_sig_in slots) for
    `validator_records` / `present_on_trusted`.  validators: [{pub_key, voting_power, address (hex)}] in set order."""
    cm = np.zeros((), COMMIT_DTYPE)
    cm["height"], cm["round"], cm["n_signatures"] = int(commit["height"]), int(commit["round"]), len(commit["signatures"])
    cm["chain_id_len"] = _put(cm["chain_id"], header["chain_id"].encode())
    bid = commit.get("block_id")
    if bid and bid.get("hash"):
        cm["has_block_id"], cm["parts_total"] = 1, int(bid["parts"]["total"])
        _put(cm["block_hash"], bytes.fromhex(bid["hash"]))
        _put(cm["parts_hash"], bytes.fromhex(bid["parts"]["hash"]))
    sg = np.zeros(n_max, COMMIT_SIG_DTYPE)
    for i, (v, cs) in enumerate(zip(validators, commit["signatures"])):
        if i >= n_max:
            break
        s = sg[i]
        _put(s["pubkey"], _b64(v["pub_key"]))
        s["voting_power"] = int(v["voting_power"])
        s["block_id_flag"] = int(cs["block_id_flag"])
        if v.get("address"):
            _put(s["address"], bytes.fromhex(v["address"]) if isinstance(v["address"], str) else bytes(v["address"]))
        if cs.get("validator_address"):
            a = cs["validator_address"]
            _put(s["sig_address"], bytes.fromhex(a) if isinstance(a, str) else bytes(a))
        if cs.get("signature"):
            _put(s["signature"], _b64(cs["signature"]))
        if cs.get("timestamp") and int(cs["block_id_flag"]) != 1:
            s["ts_seconds"], s["ts_nanos"] = parse_time(cs["timestamp"])
    return cm, sg


def header_hash(h: dict) -> bytes:
    return HeaderTree.build(header_leaves(h)).root


def inclusion_proof(tree: HeaderTree, index: int, leaf_size: int) -> np.ndarray:
    """leaf ‖ 4 aunts as one record (MerkleInclusionProofVariable, PX/frontend/merkle/tree.rs:3-8)."""
    leaf = tree.leaves[index]
    if len(leaf) != leaf_size:
        raise ValueError(f"leaf {index} is {len(leaf)} bytes, circuit expects {leaf_size}")
    return np.frombuffer(leaf + b"".join(tree.aunts(index)), dtype=np.uint8).copy()


# ------------------------------------------------------------------------------------------------
# data commitment inputs (BX/circuits/input.rs:149-271)
# ------------------------------------------------------------------------------------------------


@dataclass
class DataCommitmentInputs:
    start_header: np.ndarray          # [32]
    end_header: np.ndarray            # [32]
    dh_leaf: np.ndarray               # [B,34]
    dh_aunts: np.ndarray              # [B,4,32]
    lb_leaf: np.ndarray               # [B,72]
    lb_aunts: np.ndarray              # [B,4,32]


def get_data_commitment_inputs(trees: Dict[int, HeaderTree], start: int, end: int, max_leaves: int,
                               latest_safe: Optional[int] = None) -> DataCommitmentInputs:
    """`trees` maps block number -> HeaderTree for blocks start..end (inclusive) that exist.
    data_hash proofs cover [start, end-1], last_block_id proofs cover [start+1, end]; both are
    zero-padded to max_leaves.  If start >= end the headers are dummies (all zero)."""
    assert end - start <= max_leaves
    req_end = end if latest_safe is None else min(end, latest_safe)
    B = max_leaves
    out = DataCommitmentInputs(np.zeros(32, np.uint8), np.zeros(32, np.uint8), np.zeros((B, 34), np.uint8),
                               np.zeros((B, 4, 32), np.uint8), np.zeros((B, 72), np.uint8), np.zeros((B, 4, 32), np.uint8))
    k_dh = k_lb = 0
    for i in range(start, req_end + 1):
        t = trees[i]
        if i < req_end:
            rec = inclusion_proof(t, DATA_HASH_INDEX, 34)
            out.dh_leaf[k_dh] = rec[:34]
            out.dh_aunts[k_dh] = rec[34:].reshape(4, 32)
            k_dh += 1
        if i > start:
            rec = inclusion_proof(t, LAST_BLOCK_ID_INDEX, 72)
            out.lb_leaf[k_lb] = rec[:72]
            out.lb_aunts[k_lb] = rec[72:].reshape(4, 32)
            k_lb += 1
    if start < req_end:
        out.start_header[:] = np.frombuffer(trees[start].root, np.uint8)
        out.end_header[:] = np.frombuffer(trees[req_end].root, np.uint8)
    return out


@dataclass
class HeaderRangeMapInputs:
    """Inputs of all map jobs of one header_range proof, concatenated job-major."""
    n_jobs: int
    batch_size: int
    start_block: int
    end_block: int
    start_header: np.ndarray      # [32]
    end_header: np.ndarray        # [32]
    dh_leaf: np.ndarray           # [J*B,34]
    dh_aunts: np.ndarray          # [J*B,128]
    lb_leaf: np.ndarray           # [J*B,72]
    lb_aunts: np.ndarray          # [J*B,128]
    start_headers: np.ndarray     # [J,32]
    end_headers: np.ndarray       # [J,32]


def get_header_range_map_inputs(trees: Dict[int, HeaderTree], start: int, end: int, n_jobs: int,
                                batch_size: int, latest_safe: Optional[int] = None) -> HeaderRangeMapInputs:
    """What the 32 `DataCommitmentOffchainInputs` hints return for one range (BX/circuits/builder.rs:316-333):
    the hint of job j is called with (start+jB, start+(j+1)B) -- not with the range's end -- and clamps only to the last
    block it can fetch, latest_block - 2 (BX/circuits/input.rs:160-163) = `latest_safe`; default: the last block of the
    contiguous run start, start+1, ... present in `trees` (for a chain generated up to `end` that is `end`: the range ends
    at the chain tip)."""
    J, B = n_jobs, batch_size
    if latest_safe is None:
        latest_safe = start
        while latest_safe + 1 in trees:
            latest_safe += 1
    latest_safe = min(latest_safe, start + J * B)
    m = HeaderRangeMapInputs(J, B, start, end, np.frombuffer(trees[start].root, np.uint8).copy(),
                             np.frombuffer(trees[end].root, np.uint8).copy(), np.zeros((J * B, 34), np.uint8),
                             np.zeros((J * B, 128), np.uint8), np.zeros((J * B, 72), np.uint8),
                             np.zeros((J * B, 128), np.uint8), np.zeros((J, 32), np.uint8), np.zeros((J, 32), np.uint8))
    for j in range(J):
        bs, be = start + j * B, start + (j + 1) * B
        if bs >= latest_safe:
            continue  # nothing to fetch (start >= request_end): zero proofs, zero headers
        d = get_data_commitment_inputs(trees, bs, be, B, latest_safe=latest_safe)
        sl = slice(j * B, (j + 1) * B)
        m.dh_leaf[sl], m.dh_aunts[sl] = d.dh_leaf, d.dh_aunts.reshape(B, 128)
        m.lb_leaf[sl], m.lb_aunts[sl] = d.lb_leaf, d.lb_aunts.reshape(B, 128)
        m.start_headers[j], m.end_headers[j] = d.start_header, d.end_header
    return m


# ------------------------------------------------------------------------------------------------
# validators / votes (TX/input/conversion.rs)
# ------------------------------------------------------------------------------------------------


def validator_bytes(pubkey: bytes, power: int) -> bytes:
    """tendermint `validator.hash_bytes()`: 0a 22 0a 20 pk 10 varint(power)."""
    return b"\x0a\x22\x0a\x20" + pubkey + b"\x10" + _uvarint(power)


def vote_sign_bytes(chain_id: str, height: int, round_: int, block_id: Optional[dict], timestamp) -> bytes:
    """CanonicalVote, length-delimited (tendermint-rs SignedVote::sign_bytes)."""
    body = b"\x08\x02"
    if height:
        body += b"\x11" + int(height).to_bytes(8, "little")
    if round_:
        body += b"\x19" + int(round_).to_bytes(8, "little")
    body += _ld(0x22, _block_id(block_id))
    ts = _timestamp(timestamp)
    body += b"\x2a" + _uvarint(len(ts)) + ts
    body += _ld(0x32, chain_id.encode())
    return _uvarint(len(body)) + body


def _validator_record(pubkey: bytes, sig: bytes, msg: bytes, msg_len: int, power: int, vlen: int, signed: bool,
                      present: bool = False) -> np.ndarray:
    r = np.zeros(VAL_IN_BYTES, np.uint8)
    r[0:32] = np.frombuffer(pubkey, np.uint8)
    r[32:96] = np.frombuffer(sig, np.uint8)
    r[96:96 + len(msg)] = np.frombuffer(msg, np.uint8)
    r[220:224] = np.frombuffer(int(msg_len).to_bytes(4, "little"), np.uint8)
    r[224:232] = np.frombuffer(int(power).to_bytes(8, "little"), np.uint8)
    r[232:236] = np.frombuffer(int(vlen).to_bytes(4, "little"), np.uint8)
    r[236] = int(signed)
    r[237] = int(present)
    return r


def get_validator_data_from_block(validators: Sequence[dict], header: dict, commit: dict,
                                  n_max: int = VALIDATOR_SET_SIZE_MAX) -> np.ndarray:
    """TX/input/conversion.rs:59-140.  validators: [{pub_key(b64 or bytes), voting_power, address}]."""
    recs = []
    sigs = commit["signatures"]
    for i, cs in enumerate(sigs):
        v = validators[i]
        pk = v["pub_key"] if isinstance(v["pub_key"], (bytes, bytearray)) else base64.b64decode(v["pub_key"])
        power = int(v["voting_power"])
        vlen = len(validator_bytes(pk, power))
        if int(cs["block_id_flag"]) == 2:
            sig = cs["signature"] if isinstance(cs["signature"], (bytes, bytearray)) else base64.b64decode(cs["signature"])
            msg = vote_sign_bytes(header["chain_id"], int(commit["height"]), int(commit["round"]), commit["block_id"],
                                  cs["timestamp"])
            if len(msg) > VALIDATOR_MESSAGE_BYTES_LENGTH_MAX:
                raise ValueError("sign bytes longer than VALIDATOR_MESSAGE_BYTES_LENGTH_MAX")
            recs.append(_validator_record(pk, sig, msg, len(msg), power, vlen, True))
        else:
            recs.append(_validator_record(pk, DUMMY_SIGNATURE, b"", 32, power, vlen, False))
    for _ in range(len(sigs), n_max):
        recs.append(_validator_record(DUMMY_PUBLIC_KEY, DUMMY_SIGNATURE, b"", 32, 0, VALIDATOR_BYTE_LENGTH_MAX, False))
    return np.stack(recs)


def validator_hash_fields(validators: Sequence[dict], n_max: int = VALIDATOR_SET_SIZE_MAX):
    """TX/input/conversion.rs:142-184 -> (pubkeys [n,32], powers [n] u64, byte_lengths [n] u32)."""
    pks = np.zeros((n_max, 32), np.uint8)
    powers = np.zeros(n_max, np.uint64)
    blens = np.full(n_max, VALIDATOR_BYTE_LENGTH_MAX, np.uint32)
    pks[:] = np.frombuffer(DUMMY_PUBLIC_KEY, np.uint8)
    for i, v in enumerate(validators):
        pk = v["pub_key"] if isinstance(v["pub_key"], (bytes, bytearray)) else base64.b64decode(v["pub_key"])
        pks[i] = np.frombuffer(pk, np.uint8)
        powers[i] = int(v["voting_power"])
        blens[i] = len(validator_bytes(pk, int(v["voting_power"])))
    return pks, powers, blens


def update_present_on_trusted_header(records: np.ndarray, commit: dict, target_validators: Sequence[dict],
                                     trusted_validators: Sequence[dict]) -> None:
    """TX/input/conversion.rs:186-240 (walk the trusted set until 1/3 of target power is shared)."""
    total = sum(int(v["voting_power"]) for v in target_validators)
    tgt_idx = {v["address"]: i for i, v in enumerate(target_validators)}
    signed_addr = {s.get("validator_address") for s in commit["signatures"] if s.get("validator_address")}
    shared = 0
    for tv in trusted_validators:
        if not (total * (1.0 / 3.0) > float(shared)):
            break
        i = tgt_idx.get(tv["address"])
        if i is not None and tv["address"] in signed_addr:
            shared += int(target_validators[i]["voting_power"])
            records[i, 237] = 1
    if total * (1.0 / 3.0) > float(shared):
        raise ValueError("shared voting power is less than threshold")


def _header_common(tree: HeaderTree, header: dict, commit: dict, validators: Sequence[dict], n_max: int,
                   expected_chain_id: bytes) -> dict:
    enc_chain_id = tree.leaves[CHAIN_ID_INDEX]
    chain_buf = np.zeros(64, np.uint8)
    chain_buf[: len(enc_chain_id)] = np.frombuffer(enc_chain_id, np.uint8)
    return dict(
        n_validators=n_max,
        validators=get_validator_data_from_block(validators, header, commit, n_max),
        nb_enabled=len(validators),
        header=np.frombuffer(tree.root, np.uint8).copy(),
        height=int(header["height"]),
        round=int(commit["round"]),
        chain_id_enc=chain_buf,
        chain_id_enc_len=len(enc_chain_id),
        chain_id_aunts=np.frombuffer(b"".join(tree.aunts(CHAIN_ID_INDEX)), np.uint8).copy(),
        height_aunts=np.frombuffer(b"".join(tree.aunts(BLOCK_HEIGHT_INDEX)), np.uint8).copy(),
        height_enc_len=len(tree.leaves[BLOCK_HEIGHT_INDEX]),
        validators_hash_proof=inclusion_proof(tree, VALIDATORS_HASH_INDEX, 34),
        expected_chain_id=np.frombuffer(expected_chain_id, np.uint8).copy(),
    )


def get_skip_inputs(trusted_header: dict, trusted_validators: Sequence[dict], target_header: dict, target_commit: dict,
                    target_validators: Sequence[dict], n_max: int = VALIDATOR_SET_SIZE_MAX, skip_max: int = 100800,
                    expected_chain_id: Optional[bytes] = None) -> dict:
    """TX/input/mod.rs:426-530"""
    ttree = HeaderTree.build(header_leaves(trusted_header))
    gtree = HeaderTree.build(header_leaves(target_header))
    tgt = _header_common(gtree, target_header, target_commit, target_validators, n_max,
                         expected_chain_id or target_header["chain_id"].encode())
    update_present_on_trusted_header(tgt["validators"], target_commit, target_validators, trusted_validators)
    pks, powers, blens = validator_hash_fields(trusted_validators, n_max)
    return dict(target=tgt, trusted_block=int(trusted_header["height"]), trusted_header=np.frombuffer(ttree.root, np.uint8).copy(),
                skip_max=skip_max, trusted_validators_hash_proof=inclusion_proof(ttree, VALIDATORS_HASH_INDEX, 34),
                trusted_pubkeys=pks, trusted_powers=powers, trusted_byte_lengths=blens,
                trusted_nb_enabled=len(trusted_validators))


def get_step_inputs(prev_header: dict, next_header: dict, next_commit: dict, next_validators: Sequence[dict],
                    n_max: int = VALIDATOR_SET_SIZE_MAX, expected_chain_id: Optional[bytes] = None) -> dict:
    """TX/input/mod.rs:317-424 plus the DataCommitmentOffchainInputs<1> of prove_next_header_data_commitment."""
    ptree = HeaderTree.build(header_leaves(prev_header))
    ntree = HeaderTree.build(header_leaves(next_header))
    nxt = _header_common(ntree, next_header, next_commit, next_validators, n_max,
                         expected_chain_id or next_header["chain_id"].encode())
    return dict(next=nxt, prev_block=int(prev_header["height"]), prev_header=np.frombuffer(ptree.root, np.uint8).copy(),
                last_block_id_proof=inclusion_proof(ntree, LAST_BLOCK_ID_INDEX, 72),
                prev_next_validators_proof=inclusion_proof(ptree, NEXT_VALIDATORS_HASH_INDEX, 34),
                data_hash_proof=inclusion_proof(ptree, DATA_HASH_INDEX, 34))

"""ctypes binding of the C oracle (oracle/libbsx_oracle.so).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libbsx_oracle.so")

SUBCHAIN_BYTES = 128
SIG_OUT_BYTES = 576
VAL_IN_BYTES = 240


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_max_threads.restype = C.c_int
        for f in ("orc_sha256_pad_fixed", "orc_sha256_pad_variable", "orc_sha512_pad_variable", "orc_marshal_int64_varint",
                  "orc_marshal_validator", "orc_verify_header", "orc_verify_skip", "orc_next_header",
                  "orc_sha256_hash_input_data", "orc_sha512_hash_input_data", "orc_gate_num_constraints", "orc_gate_num_wires"):
            if hasattr(_lib, f):
                getattr(_lib, f).restype = C.c_uint32
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u8(b) -> np.ndarray:
    if isinstance(b, np.ndarray):
        return np.ascontiguousarray(b, dtype=np.uint8)
    return np.frombuffer(bytes(b), dtype=np.uint8).copy() if len(b) else np.zeros(0, np.uint8)


def sha256(msg: bytes) -> bytes:
    m = _u8(msg)
    out = np.zeros(32, np.uint8)
    lib().orc_sha256(_p(m), C.c_size_t(len(m)), _p(out))
    return out.tobytes()


def sha512(msg: bytes) -> bytes:
    m = _u8(msg)
    out = np.zeros(64, np.uint8)
    lib().orc_sha512(_p(m), C.c_size_t(len(m)), _p(out))
    return out.tobytes()


def sha256_batch(msgs: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    n = len(offsets) - 1
    out = np.zeros((n, 32), np.uint8)
    offsets = np.ascontiguousarray(offsets, np.uint32)
    lib().orc_sha256_batch(_p(_u8(msgs)), _p(offsets), C.c_uint32(n), _p(out))
    return out


def sha512_batch(msgs: np.ndarray, offsets: np.ndarray) -> np.ndarray:
    n = len(offsets) - 1
    out = np.zeros((n, 64), np.uint8)
    offsets = np.ascontiguousarray(offsets, np.uint32)
    lib().orc_sha512_batch(_p(_u8(msgs)), _p(offsets), C.c_uint32(n), _p(out))
    return out


def sha256_pad_variable(buf: bytes, length: int):
    b = _u8(buf)
    out = np.zeros(len(b) + 192, np.uint8)
    lc = C.c_uint32(0)
    n = lib().orc_sha256_pad_variable(_p(b), C.c_uint32(len(b)), C.c_uint32(length), _p(out), C.byref(lc))
    return out[: 64 * n].tobytes(), lc.value


def tm_root_from_slices(items) -> bytes:
    flat = _u8(b"".join(items))
    offs = np.zeros(len(items) + 1, np.uint32)
    offs[1:] = np.cumsum([len(i) for i in items])
    out = np.zeros(32, np.uint8)
    lib().orc_tm_root_from_slices(_p(flat), _p(offs), C.c_uint32(len(items)), _p(out))
    return out.tobytes()


def tm_aunts_from_slices(items, index: int):
    """(aunts [depth,32], root) of leaf `index` of the variable-shape Tendermint tree over `items`."""
    flat = _u8(b"".join(items))
    offs = np.zeros(len(items) + 1, np.uint32)
    offs[1:] = np.cumsum([len(i) for i in items])
    aunts, root = np.zeros((32, 32), np.uint8), np.zeros(32, np.uint8)
    lib().orc_tm_aunts_from_slices.restype = C.c_uint32
    d = lib().orc_tm_aunts_from_slices(_p(flat), _p(offs), C.c_uint32(len(items)), C.c_uint32(index), _p(aunts), _p(root))
    return aunts[:d].copy(), root.tobytes()


def header_range_inputs(n_jobs: int, B: int, headers, start: int, end: int, latest: int = None):
    """Map-circuit inputs of one range from its header records (n_jobs*B + 1 records of 512 bytes).  `latest`: the last
    block whose header can be fetched (latest_block - 2 in the reference); default = `end` (the range ends at the chain tip)."""
    latest = end if latest is None else latest
    h = _u8(headers)
    assert h.size == (n_jobs * B + 1) * 512
    slots = n_jobs * B
    out = dict(dh_leaf=np.zeros((slots, 34), np.uint8), dh_aunts=np.zeros((slots, 128), np.uint8),
               lb_leaf=np.zeros((slots, 72), np.uint8), lb_aunts=np.zeros((slots, 128), np.uint8),
               start_headers=np.zeros((n_jobs, 32), np.uint8), end_headers=np.zeros((n_jobs, 32), np.uint8),
               start_header=np.zeros(32, np.uint8), end_header=np.zeros(32, np.uint8))
    out["bad"] = lib().orc_header_range_inputs(C.c_uint32(n_jobs), C.c_uint32(B), _p(h), C.c_uint64(start), C.c_uint64(end), C.c_uint64(latest),
                                               *[_p(out[k]) for k in ("dh_leaf", "dh_aunts", "lb_leaf", "lb_aunts", "start_headers",
                                                                      "end_headers", "start_header", "end_header")])
    return out


def _field(rec, name) -> np.ndarray:
    return np.ascontiguousarray(rec[name])


def encode_header_fields(rec):
    """One decoded header (a record with the fields of include/bsx.h bsx_header_fields) -> (lens[14], fields back to back)."""
    lens, out = np.zeros(14, np.uint8), np.zeros(512, np.uint8)
    cid, lb, ph, hs, hl = (_field(rec, k) for k in ("chain_id", "last_block_hash", "parts_hash", "hashes", "hash_len"))
    f = lib().orc_encode_header_fields
    f.restype = C.c_uint32
    n = f(C.c_uint64(int(rec["version_block"])), C.c_uint64(int(rec["version_app"])), _p(cid), C.c_uint32(int(rec["chain_id_len"])),
          C.c_uint64(int(rec["height"])), C.c_int64(int(rec["time_seconds"])), C.c_uint32(int(rec["time_nanos"])),
          C.c_int(int(rec["has_last_block_id"])), _p(lb), C.c_uint32(int(rec["parts_total"])), _p(ph), _p(hs), _p(hl), _p(lens), _p(out))
    return lens, out[:n].tobytes()


def vote_sign_bytes(chain_id: bytes, height: int, round_: int, block_hash, parts_total: int, parts_hash, ts_secs: int,
                    ts_nanos: int) -> bytes:
    cid, out = _u8(chain_id), np.zeros(256, np.uint8)
    has = block_hash is not None
    bh, ph = _u8(block_hash if has else bytes(32)), _u8(parts_hash if has else bytes(32))
    f = lib().orc_vote_sign_bytes
    f.restype = C.c_uint32
    n = f(_p(cid), C.c_uint32(len(cid)), C.c_uint64(height), C.c_uint64(round_), C.c_int(int(has)), _p(bh), C.c_uint32(parts_total),
          _p(ph), C.c_int64(ts_secs), C.c_uint32(ts_nanos), _p(out))
    return out[:n].tobytes()


def validator_records(cm, sg, N: int):
    """One commit (bsx_commit_in fields) + its N signature slots (bsx_commit_sig_in fields) ->
    dict(validators [N,240], pubkeys [N,32], powers [N], byte_lengths [N], bad)."""
    n_sigs = int(cm["n_signatures"])
    k = min(n_sigs, len(sg))
    cols = {name: np.ascontiguousarray(sg[name][:k]) for name in ("pubkey", "signature", "voting_power", "ts_seconds", "ts_nanos", "block_id_flag")}
    out = dict(validators=np.zeros((N, 240), np.uint8), pubkeys=np.zeros((N, 32), np.uint8), powers=np.zeros(N, np.uint64),
               byte_lengths=np.zeros(N, np.uint32))
    cid, bh, ph = _field(cm, "chain_id"), _field(cm, "block_hash"), _field(cm, "parts_hash")
    f = lib().orc_validator_records
    f.restype = C.c_int
    out["bad"] = f(C.c_uint32(N), C.c_uint32(n_sigs), _p(cid), C.c_uint32(int(cm["chain_id_len"])), C.c_uint64(int(cm["height"])),
                   C.c_uint64(int(cm["round"])), C.c_int(int(cm["has_block_id"])), _p(bh), C.c_uint32(int(cm["parts_total"])), _p(ph),
                   _p(cols["pubkey"]), _p(cols["signature"]), _p(cols["voting_power"]), _p(cols["ts_seconds"]), _p(cols["ts_nanos"]),
                   _p(cols["block_id_flag"]), _p(out["validators"]), _p(out["pubkeys"]), _p(out["powers"]), _p(out["byte_lengths"]))
    return out


def present_on_trusted(tg, n_target: int, tr, n_trusted: int, validators: np.ndarray) -> int:
    """Marks present_on_trusted_header in validators [N,240] (in place); returns non-zero below the 1/3 threshold."""
    assert validators.dtype == np.uint8 and validators.flags.c_contiguous
    cols = [np.ascontiguousarray(tg[name][:n_target]) for name in ("address", "sig_address", "block_id_flag", "voting_power")]
    tra = np.ascontiguousarray(tr["address"][:n_trusted])
    f = lib().orc_present_on_trusted
    f.restype = C.c_int
    return f(C.c_uint32(n_target), _p(cols[0]), _p(cols[1]), _p(cols[2]), _p(cols[3]), C.c_uint32(n_trusted), _p(tra), _p(validators))


def tm_merkle_proof(leaf: bytes, aunts: bytes, depth: int, path_bits: int, hashed_leaf: bool = False):
    l, a = _u8(leaf), _u8(aunts)
    nd = 2 * depth + (0 if hashed_leaf else 1)
    dig = np.zeros((nd, 32), np.uint8)
    root = np.zeros(32, np.uint8)
    lib().orc_tm_merkle_proof(_p(l), C.c_uint32(len(l)), _p(a), C.c_uint32(depth), C.c_uint32(path_bits),
                              C.c_int(int(hashed_leaf)), _p(dig), _p(root))
    return dig, root.tobytes()


def tm_merkle_tree(leaf_digests: np.ndarray, nb_enabled: int):
    ld = _u8(leaf_digests).reshape(-1, 32)
    n = ld.shape[0]
    P = 1
    while P < n:
        P *= 2
    inner = np.zeros((P - 1, 32), np.uint8)
    root = np.zeros(32, np.uint8)
    lib().orc_tm_merkle_tree(_p(ld), C.c_uint32(n), C.c_uint64(nb_enabled), _p(inner), _p(root))
    return inner, root.tobytes()


def get_data_commitment(data_hashes: np.ndarray, start: int, end: int):
    dh = _u8(data_hashes).reshape(-1, 32)
    B = dh.shape[0]
    P = 1
    while P < B:
        P *= 2
    dig = np.zeros((B + P - 1, 32), np.uint8)
    root = np.zeros(32, np.uint8)
    fail = C.c_uint32(0)
    lib().orc_get_data_commitment(_p(dh), C.c_uint32(B), C.c_uint64(start), C.c_uint64(end), _p(dig), _p(root), C.byref(fail))
    return dig, root.tobytes(), fail.value


def prove_subchain(B, dh_leaf, dh_aunts, lb_leaf, lb_aunts, start_header, end_header, batch_start, batch_end,
                   global_end, global_end_header):
    dig = np.zeros((20 * B - 1, 32), np.uint8)
    sub = np.zeros(SUBCHAIN_BYTES, np.uint8)
    lib().orc_prove_subchain(C.c_uint32(B), _p(_u8(dh_leaf)), _p(_u8(dh_aunts)), _p(_u8(lb_leaf)), _p(_u8(lb_aunts)),
                             _p(_u8(start_header)), _p(_u8(end_header)), C.c_uint64(batch_start), C.c_uint64(batch_end),
                             C.c_uint64(global_end), _p(_u8(global_end_header)), _p(dig), _p(sub))
    return dig, sub


def prove_data_commitment(n_jobs, B, dh_leaf, dh_aunts, lb_leaf, lb_aunts, start_headers, end_headers, start_block,
                          start_header, end_block, end_header, threads=1):
    map_dig = np.zeros((n_jobs, 20 * B - 1, 32), np.uint8)
    map_sub = np.zeros((n_jobs, SUBCHAIN_BYTES), np.uint8)
    red_dig = np.zeros((n_jobs - 1, 32), np.uint8)
    red_nodes = np.zeros((n_jobs - 1, SUBCHAIN_BYTES), np.uint8)
    dc = np.zeros(32, np.uint8)
    fail = C.c_uint32(0)
    lib().orc_prove_data_commitment(
        C.c_uint32(n_jobs), C.c_uint32(B), _p(_u8(dh_leaf)), _p(_u8(dh_aunts)), _p(_u8(lb_leaf)), _p(_u8(lb_aunts)),
        _p(_u8(start_headers)), _p(_u8(end_headers)), C.c_uint64(start_block), _p(_u8(start_header)),
        C.c_uint64(end_block), _p(_u8(end_header)), _p(map_dig), _p(map_sub), _p(red_dig), _p(red_nodes), _p(dc),
        C.byref(fail), C.c_int(threads))
    return dict(map_digests=map_dig, map_subchains=map_sub, reduce_digests=red_dig, reduce_nodes=red_nodes,
                data_commitment=dc.tobytes(), fail=fail.value)


def reduce_subchains(n_jobs, B, map_subchains, start_block, start_header, end_block, end_header):
    red_dig = np.zeros((max(n_jobs - 1, 1), 32), np.uint8)
    red_nodes = np.zeros((max(n_jobs - 1, 1), SUBCHAIN_BYTES), np.uint8)
    dc = np.zeros(32, np.uint8)
    fail = C.c_uint32(0)
    lib().orc_reduce_subchains(C.c_uint32(n_jobs), C.c_uint32(B), _p(_u8(map_subchains)), C.c_uint64(start_block),
                               _p(_u8(start_header)), C.c_uint64(end_block), _p(_u8(end_header)), _p(red_dig), _p(red_nodes),
                               _p(dc), C.byref(fail))
    return dict(reduce_digests=red_dig, reduce_nodes=red_nodes, data_commitment=dc.tobytes(), fail=fail.value)


def marshal_int64_varint(v: int):
    out = np.zeros(9, np.uint8)
    n = lib().orc_marshal_int64_varint(C.c_uint64(v), _p(out))
    return out.tobytes(), n


def marshal_validator(pubkey: bytes, power: int):
    out = np.zeros(46, np.uint8)
    n = lib().orc_marshal_validator(_p(_u8(pubkey)), C.c_uint64(power), _p(out))
    return out.tobytes(), n


def hash_validator_set(pubkeys: np.ndarray, powers: np.ndarray, byte_lengths: np.ndarray, nb_enabled: int):
    pk = _u8(pubkeys).reshape(-1, 32)
    n = pk.shape[0]
    P = 1
    while P < n:
        P *= 2
    dig = np.zeros((n + P - 1, 32), np.uint8)
    root = np.zeros(32, np.uint8)
    lib().orc_hash_validator_set(C.c_uint32(n), _p(pk), _p(np.ascontiguousarray(powers, np.uint64)),
                                 _p(np.ascontiguousarray(byte_lengths, np.uint32)), C.c_uint64(nb_enabled), _p(dig), _p(root))
    return dig, root.tobytes()


def ed25519_witness(pk: bytes, sig: bytes, msg: bytes) -> bytes:
    out = np.zeros(SIG_OUT_BYTES, np.uint8)
    m = _u8(msg)
    lib().orc_ed25519_witness(_p(_u8(pk)), _p(_u8(sig)), _p(m), C.c_uint32(len(m)), _p(out))
    return out.tobytes()


def ed25519_batch(pks, sigs, msgs, msg_lens, active=None, threads=1) -> np.ndarray:
    pks = _u8(pks).reshape(-1, 32)
    n = pks.shape[0]
    out = np.zeros((n, SIG_OUT_BYTES), np.uint8)
    act = None if active is None else _u8(active)
    lib().orc_ed25519_batch(C.c_uint32(n), _p(pks), _p(_u8(sigs)), _p(_u8(msgs)), _p(np.ascontiguousarray(msg_lens, np.uint32)),
                            _p(act), _p(out), C.c_int(threads))
    return out


def ed25519_decompress(b: bytes):
    xy = np.zeros(64, np.uint8)
    root = np.zeros(32, np.uint8)
    lib().orc_ed25519_decompress.restype = C.c_int
    ok = lib().orc_ed25519_decompress(_p(_u8(b)), _p(xy), _p(root))
    return xy.tobytes(), root.tobytes(), bool(ok)


def ed25519_scalar_mul(scalar: bytes, xy: bytes, affine: bool = False) -> bytes:
    out = np.zeros(64, np.uint8)
    f = lib().orc_ed25519_affine_double_and_add if affine else lib().orc_ed25519_scalar_mul
    f(_p(_u8(scalar)), _p(_u8(xy)), _p(out))
    return out.tobytes()


def ed25519_add(a: bytes, b: bytes) -> bytes:
    out = np.zeros(64, np.uint8)
    lib().orc_ed25519_add(_p(_u8(a)), _p(_u8(b)), _p(out))
    return out.tobytes()


class _VerifyHeaderIn(C.Structure):
    _fields_ = [("n_validators", C.c_uint32), ("validators", C.c_void_p), ("nb_enabled", C.c_uint64),
                ("header", C.c_void_p), ("height", C.c_uint64), ("round", C.c_uint64), ("chain_id_enc", C.c_void_p),
                ("chain_id_enc_len", C.c_uint32), ("chain_id_aunts", C.c_void_p), ("height_aunts", C.c_void_p),
                ("height_enc_len", C.c_uint32), ("validators_hash_proof", C.c_void_p), ("expected_chain_id", C.c_void_p),
                ("expected_chain_id_len", C.c_uint32)]


class _VerifySkipIn(C.Structure):
    _fields_ = [("target", _VerifyHeaderIn), ("trusted_block", C.c_uint64), ("trusted_header", C.c_void_p),
                ("skip_max", C.c_uint32), ("trusted_validators_hash_proof", C.c_void_p), ("trusted_pubkeys", C.c_void_p),
                ("trusted_powers", C.c_void_p), ("trusted_byte_lengths", C.c_void_p), ("trusted_nb_enabled", C.c_uint64)]


class _VerifyStepIn(C.Structure):
    _fields_ = [("next", _VerifyHeaderIn), ("prev_block", C.c_uint64), ("prev_header", C.c_void_p),
                ("last_block_id_proof", C.c_void_p), ("prev_next_validators_proof", C.c_void_p),
                ("data_hash_proof", C.c_void_p)]


def _hdr_struct(h: dict, keep: list) -> _VerifyHeaderIn:
    """h: dict of numpy arrays / ints as produced by the input shapers (see tests/helpers.py)."""
    s = _VerifyHeaderIn()
    arr = lambda k: keep.append(_u8(h[k])) or keep[-1]
    s.n_validators = h["n_validators"]
    s.validators = _p(arr("validators")).value
    s.nb_enabled = h["nb_enabled"]
    s.header = _p(arr("header")).value
    s.height = h["height"]
    s.round = h["round"]
    s.chain_id_enc = _p(arr("chain_id_enc")).value
    s.chain_id_enc_len = h["chain_id_enc_len"]
    s.chain_id_aunts = _p(arr("chain_id_aunts")).value
    s.height_aunts = _p(arr("height_aunts")).value
    s.height_enc_len = h["height_enc_len"]
    s.validators_hash_proof = _p(arr("validators_hash_proof")).value
    s.expected_chain_id = _p(arr("expected_chain_id")).value
    s.expected_chain_id_len = len(h["expected_chain_id"])
    return s


def _n_header_digests(n):
    P = 1
    while P < n:
        P *= 2
    return n + P - 1 + 9 + 9 + 9


def verify_header(h: dict, threads=1):
    keep = []
    s = _hdr_struct(h, keep)
    n = h["n_validators"]
    dig = np.zeros((_n_header_digests(n), 32), np.uint8)
    ed = np.zeros((n, SIG_OUT_BYTES), np.uint8)
    fail = lib().orc_verify_header(C.byref(s), _p(dig), _p(ed), C.c_int(threads))
    return dict(sha256_digests=dig, ed=ed, fail=fail)


def verify_skip(k: dict, threads=1):
    keep = []
    s = _VerifySkipIn()
    s.target = _hdr_struct(k["target"], keep)
    n = k["target"]["n_validators"]
    arr = lambda a, dt=np.uint8: keep.append(np.ascontiguousarray(a, dtype=dt)) or keep[-1]
    s.trusted_block = k["trusted_block"]
    s.trusted_header = _p(arr(k["trusted_header"])).value
    s.skip_max = k["skip_max"]
    s.trusted_validators_hash_proof = _p(arr(k["trusted_validators_hash_proof"])).value
    s.trusted_pubkeys = _p(arr(k["trusted_pubkeys"])).value
    s.trusted_powers = _p(arr(k["trusted_powers"], np.uint64)).value
    s.trusted_byte_lengths = _p(arr(k["trusted_byte_lengths"], np.uint32)).value
    s.trusted_nb_enabled = k["trusted_nb_enabled"]
    P = 1
    while P < n:
        P *= 2
    dig = np.zeros((9 + n + P - 1 + _n_header_digests(n), 32), np.uint8)
    ed = np.zeros((n, SIG_OUT_BYTES), np.uint8)
    fail = lib().orc_verify_skip(C.byref(s), _p(dig), _p(ed), C.c_int(threads))
    return dict(sha256_digests=dig, ed=ed, fail=fail)


def next_header(k: dict, threads=1):
    keep = []
    s = _VerifyStepIn()
    s.next = _hdr_struct(k["next"], keep)
    n = k["next"]["n_validators"]
    arr = lambda a: keep.append(_u8(a)) or keep[-1]
    s.prev_block = k["prev_block"]
    s.prev_header = _p(arr(k["prev_header"])).value
    s.last_block_id_proof = _p(arr(k["last_block_id_proof"])).value
    s.prev_next_validators_proof = _p(arr(k["prev_next_validators_proof"])).value
    s.data_hash_proof = _p(arr(k["data_hash_proof"])).value
    dig = np.zeros((_n_header_digests(n) + 28, 32), np.uint8)
    ed = np.zeros((n, SIG_OUT_BYTES), np.uint8)
    dc = np.zeros(32, np.uint8)
    fail = lib().orc_next_header(C.byref(s), _p(dig), _p(ed), _p(dc), C.c_int(threads))
    return dict(sha256_digests=dig, ed=ed, data_commitment=dc.tobytes(), fail=fail)


def hash_input_data(bufs, buf_offsets, lens, kinds, sha512=False):
    """HashInputData (PX/frontend/hash/curta/mod.rs:95-192) for a request list."""
    buf_offsets = np.ascontiguousarray(buf_offsets, np.uint32)
    lens = np.ascontiguousarray(lens, np.uint32)
    kinds = np.ascontiguousarray(kinds, np.uint8)
    n = len(kinds)
    chunk = 128 if sha512 else 64
    cap = int(sum((int(buf_offsets[i + 1] - buf_offsets[i]) + 2 * chunk + chunk) // chunk + 1 for i in range(n))) + 1
    pc = np.zeros((cap, 16), np.uint64 if sha512 else np.uint32)
    eb, db, di = np.zeros(cap, np.uint8), np.zeros(cap, np.uint8), np.zeros(n, np.uint32)
    f = lib().orc_sha512_hash_input_data if sha512 else lib().orc_sha256_hash_input_data
    t = f(C.c_uint32(n), _p(_u8(bufs)), _p(buf_offsets), _p(lens), _p(kinds), _p(pc), _p(eb), _p(db), _p(di))
    return dict(padded_chunks=pc[:t], end_bits=eb[:t], digest_bits=db[:t], digest_indices=di)


GATE_U32_ARITHMETIC, GATE_U32_ADD_MANY, GATE_U32_SUBTRACTION, GATE_U32_COMPARISON, GATE_U32_RANGE_CHECK = range(5)


def gate_num_wires(gate, p0, p1=0) -> int:
    return lib().orc_gate_num_wires(C.c_uint32(gate), C.c_uint32(p0), C.c_uint32(p1))


def gate_num_constraints(gate, p0, p1=0) -> int:
    return lib().orc_gate_num_constraints(C.c_uint32(gate), C.c_uint32(p0), C.c_uint32(p1))


def gate_eval(gate, p0, p1, wires: np.ndarray, threads=1) -> np.ndarray:
    """wires [n_wires, rows] u64 (wire-major) -> constraints [n_constraints, rows]."""
    wires = np.ascontiguousarray(wires, np.uint64)
    rows = wires.shape[1]
    out = np.zeros((gate_num_constraints(gate, p0, p1), rows), np.uint64)
    rc = lib().orc_gate_eval(C.c_uint32(gate), C.c_uint32(p0), C.c_uint32(p1), _p(wires), C.c_uint32(rows), _p(out), C.c_int(threads))
    assert rc == 0
    return out


def gate_witness(gate, p0, p1, wires: np.ndarray, threads=1) -> np.ndarray:
    wires = np.ascontiguousarray(wires, np.uint64).copy()
    rc = lib().orc_gate_witness(C.c_uint32(gate), C.c_uint32(p0), C.c_uint32(p1), _p(wires), C.c_uint32(wires.shape[1]), C.c_int(threads))
    assert rc == 0
    return wires


def poseidon_permute(state) -> np.ndarray:
    s = np.ascontiguousarray(state, np.uint64).copy()
    lib().orc_poseidon_permute(_p(s))
    return s


def poseidon_batch(inputs: np.ndarray, offsets: np.ndarray, threads=1) -> np.ndarray:
    inputs = np.ascontiguousarray(inputs, np.uint64)
    offsets = np.ascontiguousarray(offsets, np.uint32)
    n = len(offsets) - 1
    out = np.zeros((n, 4), np.uint64)
    lib().orc_poseidon_batch(_p(inputs), _p(offsets), C.c_uint32(n), _p(out), C.c_int(threads))
    return out


def poseidon_hash_no_pad(inputs) -> np.ndarray:
    a = np.ascontiguousarray(inputs, np.uint64)
    return poseidon_batch(a, np.array([0, len(a)], np.uint32))[0]


def poseidon_round_constants() -> np.ndarray:
    lib().orc_poseidon_round_constants.restype = C.POINTER(C.c_uint64)
    return np.array(lib().orc_poseidon_round_constants()[:360], np.uint64)


def max_threads() -> int:
    return lib().orc_max_threads()


# ---- prover inner loops over Goldilocks (oracle/plonk.c) ----
GL_P = 2**64 - 2**32 + 1


def gl_root_of_unity(log_n: int) -> int:
    lib().orc_gl_root_of_unity.restype = C.c_uint64
    return int(lib().orc_gl_root_of_unity(C.c_uint32(log_n)))


def gl_coset_shift() -> int:
    lib().orc_gl_coset_shift.restype = C.c_uint64
    return int(lib().orc_gl_coset_shift())


def gl_ntt(x, inverse: bool = False) -> np.ndarray:
    """natural order in and out; x: [..., n] (each row transformed)"""
    a = np.ascontiguousarray(x, np.uint64).copy()
    n = a.shape[-1]
    log_n = n.bit_length() - 1
    assert 1 << log_n == n
    flat = a.reshape(-1, n)
    for row in flat:
        lib().orc_gl_ntt(_p(row), C.c_uint32(log_n), C.c_int(1 if inverse else 0))
    return a


def gl_lde(coeffs, rate_bits: int, shift: int = 0) -> np.ndarray:
    """[..., n] coefficients -> [..., n * 2^rate_bits] evaluations on shift * H_N, natural order"""
    a = np.ascontiguousarray(coeffs, np.uint64)
    n = a.shape[-1]
    log_n = n.bit_length() - 1
    flat = a.reshape(-1, n)
    out = np.zeros((flat.shape[0], n << rate_bits), np.uint64)
    for i, row in enumerate(flat):
        lib().orc_gl_lde(_p(np.ascontiguousarray(row)), C.c_uint32(log_n), C.c_uint32(rate_bits), C.c_uint64(shift), _p(out[i]))
    return out.reshape(a.shape[:-1] + (n << rate_bits,))


def gl_merkle(leaves, cap_height: int) -> np.ndarray:
    """leaves [n_leaves, width] -> digests [sum of layer sizes, 4] (layer 0 first, the cap last)"""
    a = np.ascontiguousarray(leaves, np.uint64)
    n, w = a.shape
    total, l = 0, n
    while l >= (1 << cap_height):
        total += l
        if l == 1:
            break
        l >>= 1
    out = np.zeros((total, 4), np.uint64)
    lib().orc_gl_merkle(_p(a), C.c_uint32(w), C.c_uint32(n), C.c_uint32(cap_height), _p(out))
    return out


def gl_quotient_combine(constraints, alphas, zh_inv) -> np.ndarray:
    c = np.ascontiguousarray(constraints, np.uint64)
    al = np.ascontiguousarray(alphas, np.uint64)
    z = np.ascontiguousarray(zh_inv, np.uint64)
    out = np.zeros((len(al), c.shape[1]), np.uint64)
    lib().orc_gl_quotient_combine(_p(c), C.c_uint32(c.shape[0]), C.c_uint32(c.shape[1]), _p(al), C.c_uint32(len(al)), _p(z), _p(out))
    return out


def gl_fri_fold(pairs, arity_bits: int, beta) -> np.ndarray:
    a = np.ascontiguousarray(pairs, np.uint64)
    n = a.shape[0]
    out = np.zeros((n >> arity_bits, 2), np.uint64)
    lib().orc_gl_fri_fold(_p(a), C.c_uint32(n), C.c_uint32(arity_bits), C.c_uint64(int(beta[0])), C.c_uint64(int(beta[1])), _p(out))
    return out


SHA256_TRACE_COLS = 176


def sha256_trace(chunks, end_bits, digest_bits, log_rows: int) -> np.ndarray:
    """padded chunks [n, 16] u32 (big-endian words) -> trace [176, 2^log_rows] u64"""
    c = np.ascontiguousarray(chunks, np.uint32).reshape(-1, 16)
    eb, db = np.ascontiguousarray(end_bits, np.uint8), np.ascontiguousarray(digest_bits, np.uint8)
    out = np.zeros((SHA256_TRACE_COLS, 1 << log_rows), np.uint64)
    lib().orc_sha256_trace(_p(c), _p(eb), _p(db), C.c_uint32(len(c)), C.c_uint32(log_rows), _p(out))
    return out


SHA512_TRACE_COLS = 338


def sha512_trace(chunks, end_bits, digest_bits, log_rows: int) -> np.ndarray:
    """padded chunks [n, 16] u64 (big-endian words) -> trace [338, 2^log_rows] u64"""
    c = np.ascontiguousarray(chunks, np.uint64).reshape(-1, 16)
    eb, db = np.ascontiguousarray(end_bits, np.uint8), np.ascontiguousarray(digest_bits, np.uint8)
    out = np.zeros((SHA512_TRACE_COLS, 1 << log_rows), np.uint64)
    lib().orc_sha512_trace(_p(c), _p(eb), _p(db), C.c_uint32(len(c)), C.c_uint32(log_rows), _p(out))
    return out


ED25519_TRACE_COLS = 1540


def ed25519_trace(scalars, points, log_rows: int, threads: int = 1):
    """scalars [n, 32] u8 little-endian, points [n, 64] u8 (x, y canonical, on the curve) -> (trace [1540, 2^log_rows] u64,
    results [n, 64] u8 = k * P).  The C restatement (oracle/ed25519.c); oracle/ed_trace.py is the Python-integer one."""
    sc = np.ascontiguousarray(scalars, np.uint8).reshape(-1, 32)
    pt = np.ascontiguousarray(points, np.uint8).reshape(-1, 64)
    out = np.zeros((ED25519_TRACE_COLS, 1 << log_rows), np.uint64)
    res = np.zeros((len(sc), 64), np.uint8)
    lib().orc_ed25519_trace.restype = C.c_int
    ok = lib().orc_ed25519_trace(_p(sc), _p(pt), C.c_uint32(len(sc)), C.c_uint32(log_rows), _p(res), _p(out), C.c_int(threads))
    assert ok == 1, "orc_ed25519_trace: an operation's identity did not hold (points off the curve?)"
    return out, res

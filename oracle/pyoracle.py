"""Pure-Python restatement of the Blobstream X hash / signature path (hashlib + Python ints).

TEST INFRASTRUCTURE ONLY -- never imported by the product package.  It exists to PIN the C
oracle (oracle/*.c) and, through it, the CUDA path: every function here is checked against the
reference's own fixtures / KATs in tests/test_oracle_golden.py (committed under tests/golden/).
Small cases only (pure-Python loops).

Reference paths: BX = /root/reference, TX = BX/contracts/lib/tendermintx/circuits,
PX = BX/contracts/lib/succinctx/plonky2x/core/src.
"""
from __future__ import annotations

import base64
import datetime as _dt
import hashlib
import re
from typing import Iterable, List, Sequence, Tuple

# --------------------------------------------------------------------------------------------
# protobuf field encoders (tendermint-rs 0.33.2 `encode_vec`; call sites
# TX/input/tendermint_utils.rs:374-393).  Pinned by the header-hash chain of the fixtures.
# --------------------------------------------------------------------------------------------


def varint(n: int) -> bytes:
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def pb_bytes(field: int, b: bytes) -> bytes:
    return b"" if len(b) == 0 else bytes([(field << 3) | 2]) + varint(len(b)) + b


def pb_varint(field: int, n: int) -> bytes:
    return b"" if n == 0 else bytes([(field << 3) | 0]) + varint(n & ((1 << 64) - 1))


def parse_rfc3339(ts: str) -> Tuple[int, int]:
    m = re.match(r"^(\d{4}-\d{2}-\d{2}T\d{2}:\d{2}:\d{2})(?:\.(\d+))?Z$", ts)
    assert m, ts
    secs = int(_dt.datetime.strptime(m.group(1), "%Y-%m-%dT%H:%M:%S").replace(tzinfo=_dt.timezone.utc).timestamp())
    nanos = int((m.group(2) or "0").ljust(9, "0")[:9])
    return secs, nanos


def enc_timestamp(ts: str) -> bytes:
    s, n = parse_rfc3339(ts)
    return pb_varint(1, s) + pb_varint(2, n)


def enc_block_id(bid: dict | None) -> bytes:
    if not bid or not bid.get("hash"):
        return b""
    psh = pb_varint(1, int(bid["parts"]["total"])) + pb_bytes(2, bytes.fromhex(bid["parts"]["hash"]))
    return pb_bytes(1, bytes.fromhex(bid["hash"])) + pb_bytes(2, psh)


def header_fields(h: dict) -> List[bytes]:
    """The 14 leaves, order of TX/input/tendermint_utils.rs:374-393."""
    hx = lambda k: bytes.fromhex(h.get(k) or "")
    return [
        pb_varint(1, int(h["version"]["block"])) + pb_varint(2, int(h["version"].get("app", 0) or 0)),
        pb_bytes(1, h["chain_id"].encode()),
        pb_varint(1, int(h["height"])),
        enc_timestamp(h["time"]),
        enc_block_id(h.get("last_block_id")),
        pb_bytes(1, hx("last_commit_hash")),
        pb_bytes(1, hx("data_hash")),
        pb_bytes(1, hx("validators_hash")),
        pb_bytes(1, hx("next_validators_hash")),
        pb_bytes(1, hx("consensus_hash")),
        pb_bytes(1, hx("app_hash")),
        pb_bytes(1, hx("last_results_hash")),
        pb_bytes(1, hx("evidence_hash")),
        pb_bytes(1, hx("proposer_address")),
    ]


# --------------------------------------------------------------------------------------------
# Tendermint Merkle (TX/input/tendermint_utils.rs:214-372 and PX/frontend/merkle/tendermint.rs)
# --------------------------------------------------------------------------------------------


def sha256(b: bytes) -> bytes:
    return hashlib.sha256(b).digest()


def leaf_hash(b: bytes) -> bytes:
    return sha256(b"\x00" + b)


def inner_hash(l: bytes, r: bytes) -> bytes:
    return sha256(b"\x01" + l + r)


def split_point(n: int) -> int:
    k = 1 << (n.bit_length() - 1)
    return k >> 1 if k == n else k


def tm_root(items: Sequence[bytes]) -> bytes:
    n = len(items)
    if n == 0:
        return sha256(b"")
    if n == 1:
        return leaf_hash(items[0])
    k = split_point(n)
    return inner_hash(tm_root(items[:k]), tm_root(items[k:]))


def tm_proof(items: Sequence[bytes], index: int) -> List[bytes]:
    """Aunts leaf-to-root (flatten_aunts order, tendermint_utils.rs:188-211)."""
    n = len(items)
    if n == 1:
        return []
    k = split_point(n)
    if index < k:
        return tm_proof(items[:k], index) + [tm_root(items[k:])]
    return tm_proof(items[k:], index - k) + [tm_root(items[:k])]


def header_hash(h: dict) -> bytes:
    return tm_root(header_fields(h))


def path_bits(index: int, depth: int = 4) -> List[bool]:
    """TX/builder/shared.rs:45-65"""
    return [bool((index >> i) & 1) for i in range(depth)]


def merkle_proof_schedule(leaf: bytes, aunts: Sequence[bytes], path: Sequence[bool], hashed_leaf=False):
    """PX/frontend/merkle/tendermint.rs:62-93 -> (digests in request order, root)."""
    digests = []
    h = leaf
    if not hashed_leaf:
        h = leaf_hash(leaf)
        digests.append(h)
    for a, p in zip(aunts, path):
        l, r = inner_hash(h, a), inner_hash(a, h)
        digests += [l, r]
        h = r if p else l
    return digests, h


def merkle_tree_schedule(leaf_digests: Sequence[bytes], nb_enabled: int):
    """PX/frontend/merkle/tendermint.rs:124-204 -> (raw inner digests layer-major, root)."""
    n = len(leaf_digests)
    P = 1
    while P < n:
        P *= 2
    nodes = list(leaf_digests) + [bytes(32)] * (P - n)
    en, e = [], True
    for i in range(P):
        e = e and (i != nb_enabled)
        en.append(e)
    inner = []
    while len(nodes) > 1:
        nn, ne = [], []
        for i in range(0, len(nodes), 2):
            ih = inner_hash(nodes[i], nodes[i + 1])
            inner.append(ih)
            nn.append(ih if (en[i] and en[i + 1]) else nodes[i])
            ne.append(not ((not en[i]) and (not en[i + 1])))
        nodes, en = nn, ne
    return inner, nodes[0]


def encode_data_root_tuple(data_hash: bytes, height: int) -> bytes:
    """BX/circuits/builder.rs:82-103"""
    return bytes(24) + height.to_bytes(8, "big") + data_hash


def data_commitment(data_hashes: Sequence[bytes], start: int) -> bytes:
    """Celestia data commitment over [start, start+len): variable-shape tree of tuples."""
    return tm_root([encode_data_root_tuple(d, start + i) for i, d in enumerate(data_hashes)])


# --------------------------------------------------------------------------------------------
# validators (TX/builder/validator.rs:185-252, TX/builder/shared.rs:67-156)
# --------------------------------------------------------------------------------------------


def marshal_int64_varint9(v: int) -> Tuple[bytes, int]:
    sept = [(v >> (7 * i)) & 0x7F for i in range(9)]
    last = max([i for i, s in enumerate(sept) if s] or [0])
    return bytes(s | (0x80 if i < last else 0) for i, s in enumerate(sept)), last + 1


def marshal_validator(pubkey: bytes, power: int) -> bytes:
    """Simple (unpadded) validator bytes: 0a 22 0a 20 pk 10 varint."""
    return b"\x0a\x22\x0a\x20" + pubkey + b"\x10" + varint(power)


def validators_hash(vals: Iterable[Tuple[bytes, int]]) -> bytes:
    return tm_root([marshal_validator(pk, p) for pk, p in vals])


# --------------------------------------------------------------------------------------------
# CanonicalVote sign-bytes (tendermint-rs SignedVote::sign_bytes; TX/input/conversion.rs:34-39)
# --------------------------------------------------------------------------------------------


def canonical_vote_sign_bytes(chain_id: str, height: int, round_: int, block_id: dict | None, timestamp: str) -> bytes:
    body = pb_varint(1, 2)  # SignedMsgType::Precommit
    if height:
        body += b"\x11" + height.to_bytes(8, "little")
    if round_:
        body += b"\x19" + round_.to_bytes(8, "little")
    if block_id and block_id.get("hash"):
        psh = pb_varint(1, int(block_id["parts"]["total"])) + pb_bytes(2, bytes.fromhex(block_id["parts"]["hash"]))
        cbid = pb_bytes(1, bytes.fromhex(block_id["hash"])) + pb_bytes(2, psh)
        body += pb_bytes(4, cbid)
    ts = enc_timestamp(timestamp)
    body += b"\x2a" + varint(len(ts)) + ts
    body += pb_bytes(6, chain_id.encode())
    return varint(len(body)) + body


# --------------------------------------------------------------------------------------------
# Ed25519 with Python ints (starkyx AffinePoint / decompress semantics)
# --------------------------------------------------------------------------------------------

P25519 = 2**255 - 19
L25519 = 2**252 + 27742317777372353535851937790883648493
D25519 = (-121665 * pow(121666, P25519 - 2, P25519)) % P25519
SQRT_M1 = pow(2, (P25519 - 1) // 4, P25519)
GY = 4 * pow(5, P25519 - 2, P25519) % P25519


def ed_decompress(b: bytes):
    """-> ((x, y), root, ok): root = even sqrt of (y^2-1)/(d y^2+1); x = root or p-root by sign."""
    n = int.from_bytes(b, "little")
    sign = n >> 255
    y = (n & ((1 << 255) - 1)) % P25519
    u = (y * y - 1) % P25519
    v = (D25519 * y * y + 1) % P25519
    x2 = u * pow(v, P25519 - 2, P25519) % P25519
    beta = pow(x2, (P25519 + 3) // 8, P25519)
    if (beta * beta) % P25519 == (-x2) % P25519 and x2 != 0:
        beta = beta * SQRT_M1 % P25519
    ok = (beta * beta) % P25519 == x2
    if not ok:  # the reference panics; shared convention: identity point, root 0
        return (0, 1), 0, False
    if beta & 1:
        beta = (P25519 - beta) % P25519
    x = (P25519 - beta) % P25519 if sign else beta
    return (x, y), beta, ok


GX = ed_decompress(GY.to_bytes(32, "little"))[0][0]
G = (GX, GY)


def ed_add(p, q):
    x1, y1 = p
    x2, y2 = q
    t = D25519 * x1 * x2 * y1 * y2 % P25519
    x3 = (x1 * y2 + x2 * y1) * pow(1 + t, P25519 - 2, P25519) % P25519
    y3 = (y1 * y2 + x1 * x2) * pow(1 - t, P25519 - 2, P25519) % P25519
    return (x3, y3)


def ed_mul(k: int, p):
    # extended coordinates to keep it quick
    def ext_add(P, Q):
        X1, Y1, Z1, T1 = P
        X2, Y2, Z2, T2 = Q
        A = (Y1 - X1) * (Y2 - X2) % P25519
        B = (Y1 + X1) * (Y2 + X2) % P25519
        C = T1 * 2 * D25519 * T2 % P25519
        Dd = Z1 * 2 * Z2 % P25519
        E, F, Gg, H = B - A, Dd - C, Dd + C, B + A
        return (E * F % P25519, Gg * H % P25519, F * Gg % P25519, E * H % P25519)

    Q = (0, 1, 1, 0)
    R = (p[0], p[1], 1, p[0] * p[1] % P25519)
    while k:
        if k & 1:
            Q = ext_add(Q, R)
        R = ext_add(R, R)
        k >>= 1
    zi = pow(Q[2], P25519 - 2, P25519)
    return (Q[0] * zi % P25519, Q[1] * zi % P25519)


def ed_point_bytes(p) -> bytes:
    return p[0].to_bytes(32, "little") + p[1].to_bytes(32, "little")


def ed_witness(pk: bytes, sig: bytes, msg: bytes) -> dict:
    """PX/frontend/ecc/curve25519/ed25519/eddsa.rs:161-203, one signature."""
    digest = hashlib.sha512(sig[:32] + pk + msg).digest()
    hi = int.from_bytes(digest, "little")
    div, h = divmod(hi, L25519)
    s = int.from_bytes(sig[32:], "little")
    sG = ed_mul(s, G)
    A, a_root, a_ok = ed_decompress(pk)
    hA = ed_mul(h, A)
    R, r_root, r_ok = ed_decompress(sig[:32])
    total = ed_add(R, hA)
    return dict(digest=digest, h=h, div=div, sG=sG, A=A, A_root=a_root, hA=hA, R=R, R_root=r_root, sum=total,
                s_lt_l=s < L25519, a_ok=a_ok, r_ok=r_ok, verified=(sG == total))


def ed_witness_bytes(pk: bytes, sig: bytes, msg: bytes) -> bytes:
    """Same record layout as ORC_SIG_OUT_BYTES in bsx_oracle.h."""
    w = ed_witness(pk, sig, msg)
    out = bytearray(576)
    out[0:64] = w["digest"]
    out[64:96] = w["h"].to_bytes(32, "little")
    out[96:136] = w["div"].to_bytes(40, "little")
    out[136:200] = ed_point_bytes(w["sG"])
    out[200:264] = ed_point_bytes(w["A"])
    out[264:296] = w["A_root"].to_bytes(32, "little")
    out[296:360] = ed_point_bytes(w["hA"])
    out[360:424] = ed_point_bytes(w["R"])
    out[424:456] = w["R_root"].to_bytes(32, "little")
    out[456:520] = ed_point_bytes(w["sum"])
    out[520] = (1 if w["s_lt_l"] else 0) | (2 if w["a_ok"] else 0) | (4 if w["r_ok"] else 0) | (8 if w["verified"] else 0)
    return bytes(out)


DUMMY_PUBLIC_KEY = bytes([138, 136, 227, 221, 116, 9, 241, 149, 253, 82, 219, 45, 60, 186, 93, 114, 202, 103, 9, 191,
                          29, 148, 18, 27, 243, 116, 136, 1, 180, 15, 111, 92])
DUMMY_SIGNATURE = bytes([55, 20, 104, 158, 84, 120, 194, 17, 6, 237, 157, 164, 85, 88, 158, 137, 187, 119, 187, 240,
                         159, 73, 80, 63, 133, 162, 74, 91, 48, 53, 6, 138, 1, 41, 22, 121, 249, 46, 198, 145, 155, 102,
                         3, 210, 168, 135, 173, 55, 252, 72, 45, 126, 169, 178, 191, 7, 153, 67, 112, 90, 150, 33, 140, 7])


def b64(s: str) -> bytes:
    return base64.b64decode(s)

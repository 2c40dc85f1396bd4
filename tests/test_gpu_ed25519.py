"""GPU parity for the Ed25519 witness kernel (K4+K5) through the C ABI, bit-exact vs the oracle."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
L = 2**252 + 27742317777372353535851937790883648493


@pytest.fixture(scope="module")
def ctx():
    from blobstreamx_b200 import lib
    c = lib.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def orc():
    from oracle import cbind
    return cbind


@pytest.mark.parametrize("n", [1, 100, 333])
def test_ed25519_batch_synthetic(ctx, orc, n):
    """config 5 shape: CanonicalVote messages padded to 124, ~1% inactive (DUMMY) lanes."""
    from blobstreamx_b200 import synthetic as S
    pks, sigs, msgs, lens, active = S.ed25519_batch_inputs(n, inactive_every=7)
    got = ctx.ed25519_batch(pks, sigs, msgs, lens, active)
    want = orc.ed25519_batch(pks, sigs, msgs, lens, active, threads=8)
    assert (got == want).all()
    assert (got[:, 520] == 0xF).all()  # every lane verifies (inactive lanes on the DUMMY triple)
    # no active mask / no lens (fixed-length messages)
    got2 = ctx.ed25519_batch(pks, sigs, msgs, None, None)
    want2 = orc.ed25519_batch(pks, sigs, msgs, np.full(n, 124, np.uint32), None, threads=8)
    assert (got2 == want2).all()


def test_ed25519_large_batch_one_thread_per_signature_path(ctx, orc):
    """Above 16 384 signatures bsx_ed25519_batch switches from the three-stage quad-lane kernels to the
    one-thread-per-signature kernel (k_ed25519.cu); both must produce the oracle's records, including DUMMY lanes
    and a corrupted signature in the middle of the batch."""
    from blobstreamx_b200 import synthetic as S
    base = S.ed25519_batch_inputs(700, inactive_every=9)
    n = 16384 + 516
    idx = np.arange(n) % 700
    pks, sigs, msgs, lens, active = (np.ascontiguousarray(a[idx]) for a in base)
    sigs[9000, 3] ^= 0x10
    got = ctx.ed25519_batch(pks, sigs, msgs, lens, active)
    want = orc.ed25519_batch(pks, sigs, msgs, lens, active, threads=8)
    assert (got == want).all()
    assert (np.delete(got[:, 520], 9000) == 0xF).all() and got[9000, 520] & 8 == 0
    # the same inputs cut below the threshold run on the quad-lane path and must agree record for record
    small = ctx.ed25519_batch(pks[:9100], sigs[:9100], msgs[:9100], lens[:9100], active[:9100])
    assert (small == got[:9100]).all()


def test_ed25519_negative_and_edge_cases(ctx, orc):
    """Flipped bits must not verify (eddsa.rs:344-386 must-panic test); s >= l; undecodable points;
    random garbage -- flags and every intermediate value equal the oracle's."""
    from nacl.signing import SigningKey
    rng = np.random.default_rng(11)
    pk_l, sig_l, msg_l = [], [], []
    for i in range(16):
        sk = SigningKey(hashlib.sha256(b"neg%d" % i).digest())
        m = rng.bytes(int(rng.integers(0, 125)))
        pk_l.append(bytes(sk.verify_key)); sig_l.append(sk.sign(m).signature); msg_l.append(m)
    pk, sig, m = pk_l[0], sig_l[0], msg_l[0]
    extra = [(pk, sig[:5] + bytes([sig[5] ^ 1]) + sig[6:], m), (pk, sig[:40] + bytes([sig[40] ^ 4]) + sig[41:], m),
             (pk, sig, (m + b"x")[:124]), (pk, sig[:32] + (2**256 - 1).to_bytes(32, "little"), m),
             (pk, sig[:32] + L.to_bytes(32, "little"), m), ((2).to_bytes(32, "little"), sig, m),
             (pk, (2).to_bytes(32, "little") + sig[32:], m), (bytes(32), bytes(64), b""),
             ((1).to_bytes(32, "little"), (1).to_bytes(32, "little") + bytes(32), b"")]
    extra += [(rng.bytes(32), rng.bytes(64), rng.bytes(60)) for _ in range(23)]
    for a, b, c in extra:
        pk_l.append(a); sig_l.append(b); msg_l.append(c)
    n = len(pk_l)
    pks = np.frombuffer(b"".join(pk_l), np.uint8).reshape(n, 32)
    sigs = np.frombuffer(b"".join(sig_l), np.uint8).reshape(n, 64)
    msgs = np.zeros((n, 124), np.uint8)
    lens = np.zeros(n, np.uint32)
    for i, mm in enumerate(msg_l):
        msgs[i, :len(mm)] = np.frombuffer(mm, np.uint8)
        lens[i] = len(mm)
    got = ctx.ed25519_batch(pks, sigs, msgs, lens)
    want = orc.ed25519_batch(pks, sigs, msgs, lens, None, threads=8)
    assert (got == want).all()
    flags = got[:, 520]
    assert (flags[:16] == 0xF).all() and (flags[16:19] & 8 == 0).all() and flags[19] & 1 == 0 and flags[21] & 2 == 0


@pytest.mark.parametrize("height", ["10000", "157001"])
def test_ed25519_fixture_commits(ctx, orc, golden, height):
    """Real mocha-4 commits (TX fixtures; 157001 = 100 validators, 98 commit sigs, 1 nil, 1 absent):
    every lane (real or DUMMY) verifies on the GPU and all records equal the oracle's."""
    from blobstreamx_b200 import inputs as I
    hdr, commit, vals = golden["headers"][height], golden["commits"][height], golden["validators"][height]
    recs = I.get_validator_data_from_block(vals, hdr, commit, 100)
    lens = recs[:, 220:224].copy().view(np.uint32).reshape(-1)
    a = (recs[:, 0:32].copy(), recs[:, 32:96].copy(), recs[:, 96:220].copy(), lens, recs[:, 236].copy())
    got = ctx.ed25519_batch(*a)
    assert (got[:, 520] == 0xF).all()
    assert (got == orc.ed25519_batch(*a, threads=8)).all()

"""Stress inputs for the Ed25519 witness kernels: valid signatures, a third of them corrupted in one bit of pk / R / s /
message, random garbage, and edge encodings (y = 0, 1, p-1, p, 2^255-1, a small-order point, s around l and 2^256).
Shared by tests/test_gpu_ed25519_builds.py and scripts/stress_ed25519.py."""
import numpy as np

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493


def stress_inputs(n_valid: int, n_garb: int, seed: int = 20261017, base_n: int = 2000):
    from blobstreamx_b200 import synthetic as S
    rng = np.random.default_rng(seed)
    base = S.ed25519_batch_inputs(base_n, inactive_every=50)
    idx = rng.integers(0, base_n, n_valid)
    pks, sigs, msgs, lens, act = (np.ascontiguousarray(a[idx]) for a in base)
    for i in range(0, n_valid, 3):
        which = rng.integers(0, 4)
        tgt = (pks, sigs, sigs, msgs)[which]
        lo, hi = ((0, 32), (0, 32), (32, 64), (0, max(1, int(lens[i]))))[which]
        tgt[i, rng.integers(lo, hi)] ^= np.uint8(1 << rng.integers(0, 8))
    g_pk = rng.integers(0, 256, (n_garb, 32), dtype=np.uint8)
    g_sig = rng.integers(0, 256, (n_garb, 64), dtype=np.uint8)
    g_msg = rng.integers(0, 256, (n_garb, 124), dtype=np.uint8)
    g_len = rng.integers(0, 125, n_garb).astype(np.uint32)
    edge_y = [0, 1, 2, P - 1, P, P + 1, 2**255 - 1, 2**255 - 19 + 18, 19, 2**254, (1 << 255) | 1, (1 << 255), 2**256 - 1,
              0x7a03ac9277fdc74ec6cc392cfa53202a0f67100d760b3cba4fd84d3d706a17c7]   # a small-order y
    edge_s = [0, 1, L - 1, L, L + 1, 2**252, 2**253 - 1, 2**256 - 1]
    for j in range(min(n_garb, 4000)):
        if j % 2 == 0:
            g_pk[j] = np.frombuffer((edge_y[(j // 2) % len(edge_y)] % 2**256).to_bytes(32, "little"), np.uint8)
        else:
            g_sig[j, :32] = np.frombuffer((edge_y[(j // 2) % len(edge_y)] % 2**256).to_bytes(32, "little"), np.uint8)
        g_sig[j, 32:] = np.frombuffer(edge_s[j % len(edge_s)].to_bytes(32, "little"), np.uint8)
    # edges of the R shortcut of the FP64 witness core (R' = sG - hA compared with the encoded R): A = identity (or undecodable,
    # which falls back to it) and s = 0 make R' the identity; R encoded canonically, with the sign bit on x = 0, as y = 1 + p
    k = n_garb - 1
    for pk_v in (1, 2):
        for r_v in (1, 1 | 1 << 255, 1 + P, 1 + P | 1 << 255):
            if k < 0:
                break
            g_pk[k] = np.frombuffer(pk_v.to_bytes(32, "little"), np.uint8)
            g_sig[k, :32] = np.frombuffer(r_v.to_bytes(32, "little"), np.uint8)
            g_sig[k, 32:] = 0
            k -= 1
    if n_valid > 1:
        sigs[1, 31] ^= 0x80     # -R of a signature: decodes, does not verify (entry 1 is not one of the corrupted ones)
    return (np.concatenate([pks, g_pk]), np.concatenate([sigs, g_sig]), np.concatenate([msgs, g_msg]),
            np.concatenate([lens, g_len]), np.concatenate([act, np.ones(n_garb, np.uint8)]))

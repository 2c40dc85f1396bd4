"""Synthetic inputs of the device-side encoders (shared by the CPU and GPU tests): random decoded headers, commits with
mixed vote kinds and overlapping trusted / target validator sets."""
import numpy as np

from blobstreamx_b200 import inputs as I


def random_header_fields(n: int, seed: int = 11) -> np.ndarray:
    rng = np.random.default_rng(seed)
    f = np.zeros(n, I.HEADER_FIELDS_DTYPE)
    for i, r in enumerate(f):
        r["version_block"] = int(rng.integers(0, 3)) * 11
        r["version_app"] = int(rng.integers(0, 3))
        r["height"] = 0 if i == 0 else int(rng.integers(1, 1 << int(rng.integers(1, 63))))
        r["time_seconds"] = 0 if i == 1 else int(rng.integers(1, 1 << 33))
        r["time_nanos"] = 0 if i % 5 == 2 else int(rng.integers(1, 10**9))
        cl = 56 if i == 4 else int(rng.integers(0, 51))       # 56 = the whole array
        r["chain_id_len"] = cl
        r["chain_id"][:cl] = rng.integers(97, 123, cl)
        r["has_last_block_id"] = int(i % 7 != 3)
        r["parts_total"] = int(rng.integers(0, 3)) * int(rng.integers(1, 1 << 20))
        r["last_block_hash"] = rng.integers(0, 256, 32)
        r["parts_hash"] = rng.integers(0, 256, 32)
        for k in range(9):
            ln = 20 if k == 8 else int(rng.choice([0, 32, 32, 32, 8]))
            r["hash_len"][k] = ln
            r["hashes"][k][:ln] = rng.integers(0, 256, ln)
    return f


def chain_header_fields(n: int, seed: int = 13) -> np.ndarray:
    """Headers as a live chain has them: one chain id, consecutive heights, 12 s block time, every hash present."""
    rng = np.random.default_rng(seed)
    f = np.zeros(n, I.HEADER_FIELDS_DTYPE)
    f["version_block"], f["version_app"] = 11, 1
    f["height"] = 1_000_000 + np.arange(n)
    f["time_seconds"] = 1_700_000_000 + 12 * np.arange(n)
    f["time_nanos"] = rng.integers(0, 10**9, n)
    cid = b"celestia"
    f["chain_id_len"] = len(cid)
    f["chain_id"][:, : len(cid)] = np.frombuffer(cid, np.uint8)
    f["has_last_block_id"], f["parts_total"] = 1, 1
    f["last_block_hash"] = rng.integers(0, 256, (n, 32))
    f["parts_hash"] = rng.integers(0, 256, (n, 32))
    f["hash_len"][:, :8], f["hash_len"][:, 8] = 32, 20
    f["hashes"] = rng.integers(0, 256, (n, 9, 32))
    f["hashes"][:, 8, 20:] = 0
    return f


def random_commits(n: int, N: int, seed: int = 12):
    """-> commits [n], target slots [n, N], trusted slots [n, N], n_target [n], n_trusted [n]"""
    rng = np.random.default_rng(seed)
    cm = np.zeros(n, I.COMMIT_DTYPE)
    tg = np.zeros((n, N), I.COMMIT_SIG_DTYPE)
    tr = np.zeros((n, N), I.COMMIT_SIG_DTYPE)
    n_tr = np.zeros(n, np.uint32)
    for c in range(n):
        k = int(rng.integers(1, N + 1))
        if c == 2:
            k = N + 3                                   # set larger than VALIDATOR_SET_SIZE_MAX: flagged
        cm[c]["height"] = 0 if c == 3 else int(rng.integers(1, 1 << 40))
        cm[c]["round"] = int(rng.integers(0, 3))
        cm[c]["n_signatures"] = k
        cm[c]["block_hash"] = rng.integers(0, 256, 32)
        cm[c]["parts_hash"] = rng.integers(0, 256, 32)
        cm[c]["parts_total"] = int(rng.integers(0, 4))
        cl = 50 if c == 4 else int(rng.integers(0, 9))  # a 50-byte chain id pushes the message past 124 bytes
        cm[c]["chain_id_len"] = cl
        cm[c]["chain_id"][:cl] = rng.integers(97, 123, cl)
        cm[c]["has_block_id"] = int(c % 9 != 5)
        kk = min(k, N)
        s = tg[c]
        s["pubkey"][:kk] = rng.integers(0, 256, (kk, 32))
        s["signature"][:kk] = rng.integers(0, 256, (kk, 64))
        s["voting_power"][:kk] = rng.integers(1, 1 << 40, kk).astype(np.uint64) >> rng.integers(0, 36, kk).astype(np.uint64)
        s["voting_power"][:kk] = np.maximum(s["voting_power"][:kk], 1)
        s["ts_seconds"][:kk] = rng.integers(0, 1 << 33, kk) * (rng.random(kk) > 0.05)
        s["ts_nanos"][:kk] = rng.integers(0, 10**9, kk) * (rng.random(kk) > 0.1)
        s["block_id_flag"][:kk] = rng.choice([1, 2, 2, 2, 2, 3], kk)
        s["address"][:kk] = rng.integers(0, 256, (kk, 20))
        s["sig_address"][:kk] = s["address"][:kk]
        # trusted set: a shuffled mix of target validators and strangers
        m = int(rng.integers(1, N + 1))
        n_tr[c] = m
        share = rng.random() if c % 4 else 0.02          # some commits fall short of a third
        for j in range(m):
            if rng.random() < share:
                tr[c][j]["address"] = s["address"][int(rng.integers(0, kk))]
            else:
                tr[c][j]["address"] = rng.integers(0, 256, 20)
    return cm, tg, tr, np.minimum(cm["n_signatures"], N).astype(np.uint32), n_tr

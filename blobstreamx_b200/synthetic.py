"""Synthetic Celestia-like chains for the BASELINE configs.

Deterministic from a seed: header i has 14 protobuf leaves with chain_id "celestia",
height = start+i, time = epoch + 12 s * i, last_block_id.hash = hash(header i-1) (the chain links),
random 32-byte data/app/... hashes, and a fixed 100-validator set whose Ed25519 keys are
sk_j = SHA256(seed ‖ j); every validator signs the target header in round 0.
This is workload generation (host side, hashlib / PyNaCl); nothing here is on the measured path.
"""
from __future__ import annotations

import hashlib
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import inputs as I

SEED = 0xB10B5
EPOCH = 1_700_000_000
CHAIN_ID = "celestia"


def _rng(seed: int, stream: int) -> np.random.Generator:
    return np.random.Generator(np.random.PCG64([seed, stream]))


@dataclass
class ValidatorSet:
    secret_keys: List[bytes]
    validators: List[dict]          # [{address, pub_key(bytes), voting_power}]
    hash: bytes

    @staticmethod
    def make(seed: int = SEED, n: int = 100) -> "ValidatorSet":
        from nacl.signing import SigningKey

        rng = _rng(seed, 1)
        powers = rng.integers(1, 1 << 40, n, dtype=np.int64)
        sks, vals = [], []
        for j in range(n):
            sk = hashlib.sha256(seed.to_bytes(8, "little") + j.to_bytes(4, "little")).digest()
            pk = bytes(SigningKey(sk).verify_key)
            sks.append(sk)
            vals.append({"address": hashlib.sha256(pk).digest()[:20].hex().upper(), "pub_key": pk,
                         "voting_power": int(powers[j])})
        # tendermint orders a validator set by voting power (desc) then address
        order = sorted(range(n), key=lambda j: (-vals[j]["voting_power"], vals[j]["address"]))
        sks = [sks[j] for j in order]
        vals = [vals[j] for j in order]
        leaves = [I.validator_bytes(v["pub_key"], v["voting_power"]) for v in vals]
        return ValidatorSet(sks, vals, _tm_root(leaves))


def _tm_root(items: List[bytes]) -> bytes:
    n = len(items)
    if n == 0:
        return hashlib.sha256(b"").digest()
    if n == 1:
        return hashlib.sha256(b"\x00" + items[0]).digest()
    k = 1 << ((n - 1).bit_length() - 1)
    return hashlib.sha256(b"\x01" + _tm_root(items[:k]) + _tm_root(items[k:])).digest()


@dataclass
class Chain:
    start: int
    headers: List[dict]
    trees: Dict[int, I.HeaderTree]
    valset: ValidatorSet

    def tree(self, height: int) -> I.HeaderTree:
        return self.trees[height]


def make_chain(n_headers: int, start: int = 1_000_000, seed: int = SEED, valset: Optional[ValidatorSet] = None,
               chain_id: str = CHAIN_ID) -> Chain:
    """Headers start .. start+n_headers-1, hash-linked."""
    valset = valset or ValidatorSet.make(seed)
    rng = _rng(seed, 2 + start)
    rnd = rng.integers(0, 256, (n_headers, 8, 32), dtype=np.uint8)
    prop = rng.integers(0, 256, (n_headers, 20), dtype=np.uint8)
    headers, trees = [], {}
    prev_hash = hashlib.sha256(b"genesis" + seed.to_bytes(8, "little")).digest()
    vh = valset.hash.hex().upper()
    for i in range(n_headers):
        hx = lambda k: rnd[i, k].tobytes().hex().upper()
        h = {
            "version": {"block": "11", "app": "1"}, "chain_id": chain_id, "height": str(start + i),
            "time": (EPOCH + 12 * i, 0),
            "last_block_id": {"hash": prev_hash.hex().upper(), "parts": {"total": 1, "hash": hx(0)}},
            "last_commit_hash": hx(1), "data_hash": hx(2), "validators_hash": vh, "next_validators_hash": vh,
            "consensus_hash": hx(3), "app_hash": hx(4), "last_results_hash": hx(5), "evidence_hash": hx(6),
            "proposer_address": prop[i].tobytes().hex().upper(),
        }
        t = I.HeaderTree.build(I.header_leaves(h))
        headers.append(h)
        trees[start + i] = t
        prev_hash = t.root
    return Chain(start, headers, trees, valset)


def make_commit(chain: Chain, height: int, round_: int = 0, absent: Tuple[int, ...] = (), nil: Tuple[int, ...] = (),
                seed: int = SEED) -> dict:
    """All validators precommit `height` (except `absent` / `nil` indices)."""
    from nacl.signing import SigningKey

    h = chain.headers[height - chain.start]
    rng = _rng(seed, 7 + height)
    block_id = {"hash": chain.trees[height].root.hex().upper(),
                "parts": {"total": 1, "hash": rng.integers(0, 256, 32, dtype=np.uint8).tobytes().hex().upper()}}
    sigs = []
    for j, v in enumerate(chain.valset.validators):
        ts = (EPOCH + 12 * (height - chain.start) + 6, int(rng.integers(0, 10**9)))
        if j in absent:
            sigs.append({"block_id_flag": 1, "validator_address": "", "timestamp": ts, "signature": None})
            continue
        bid = None if j in nil else block_id
        msg = I.vote_sign_bytes(h["chain_id"], height, round_, bid, ts)
        sig = SigningKey(chain.valset.secret_keys[j]).sign(msg).signature
        sigs.append({"block_id_flag": 3 if j in nil else 2, "validator_address": v["address"], "timestamp": ts,
                     "signature": sig})
    return {"height": str(height), "round": round_, "block_id": block_id, "signatures": sigs}


def header_range_inputs(n_jobs: int, batch_size: int, n_blocks: Optional[int] = None, start: int = 1_000_000,
                        seed: int = SEED, valset: Optional[ValidatorSet] = None, with_skip: bool = True, extra_blocks: int = 0):
    """One header_range instance: trusted block = start, target = start + n_blocks
    (default: the full range n_jobs*batch_size).  `extra_blocks`: the chain continues that many blocks past the target
    (the map jobs past the range's end then carry real proofs, as the reference's hints fetch them).
    Returns (map_inputs, skip_inputs or None, chain)."""
    n_blocks = n_jobs * batch_size if n_blocks is None else n_blocks
    chain = make_chain(n_blocks + 1 + extra_blocks, start, seed, valset)
    m = I.get_header_range_map_inputs(chain.trees, start, start + n_blocks, n_jobs, batch_size)
    skip = None
    if with_skip:
        target = start + n_blocks
        commit = make_commit(chain, target, seed=seed)
        skip = I.get_skip_inputs(chain.headers[0], chain.valset.validators, chain.headers[n_blocks], commit,
                                 chain.valset.validators, expected_chain_id=chain.headers[0]["chain_id"].encode())
    return m, skip, chain


def data_commitment_sweep_inputs(n_trees: int, n_leaves: int = 2048, seed: int = SEED):
    """Config 4: T independent trees of `n_leaves` random data roots, sequential heights."""
    rng = _rng(seed, 11)
    data_hashes = rng.integers(0, 256, (n_trees, n_leaves, 32), dtype=np.uint8)
    starts = (1_000_000 + np.arange(n_trees, dtype=np.uint64) * n_leaves).astype(np.uint64)
    ends = starts + np.uint64(n_leaves)
    return data_hashes, starts, ends


def ed25519_batch_inputs(n: int, seed: int = SEED, inactive_every: int = 100):
    """Config 5: n CanonicalVote-shaped messages (108-109 B padded to 124), ~1% inactive lanes."""
    from nacl.signing import SigningKey

    rng = _rng(seed, 13)
    pks = np.zeros((n, 32), np.uint8)
    sigs = np.zeros((n, 64), np.uint8)
    msgs = np.zeros((n, 124), np.uint8)
    lens = np.zeros(n, np.uint32)
    active = np.ones(n, np.uint8)
    keys = [SigningKey(hashlib.sha256(seed.to_bytes(8, "little") + j.to_bytes(4, "little")).digest()) for j in range(min(n, 100))]
    for i in range(n):
        sk = keys[i % len(keys)]
        bid = {"hash": rng.integers(0, 256, 32, dtype=np.uint8).tobytes().hex(),
               "parts": {"total": 1, "hash": rng.integers(0, 256, 32, dtype=np.uint8).tobytes().hex()}}
        m = I.vote_sign_bytes(CHAIN_ID, 1_000_000 + i, 0, bid, (EPOCH + i, int(rng.integers(0, 10**9))))
        pks[i] = np.frombuffer(bytes(sk.verify_key), np.uint8)
        sigs[i] = np.frombuffer(sk.sign(m).signature, np.uint8)
        msgs[i, : len(m)] = np.frombuffer(m, np.uint8)
        lens[i] = len(m)
        if inactive_every and i % inactive_every == inactive_every - 1:
            active[i] = 0
    return pks, sigs, msgs, lens, active

"""ctypes loader for libbsx.so (the C ABI in include/bsx.h).

There is NO CPU fallback: if the CUDA library is missing or no device is present every call
raises.  PyTorch is not needed here; callers that want device-resident buffers pass raw device
pointers (e.g. tensor.data_ptr()) and a stream handle to the `*_dev` entry points.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# BSX_LIB_PATH selects an alternative build of the SAME CUDA library (kernel-variant experiments)
LIB_PATH = os.environ.get("BSX_LIB_PATH") or os.path.join(_HERE, "libbsx.so")

SUBCHAIN_BYTES = 128
SIG_OUT_BYTES = 576
VAL_IN_BYTES = 240


class BsxError(RuntimeError):
    pass


_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BsxError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.bsx_last_error.restype = C.c_char_p
        _lib.bsx_launch_count.restype = C.c_uint64
        _lib.bsx_init.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
        _lib.bsx_verify_digest_count.restype = C.c_uint32
        _lib.bsx_gate_num_wires.restype = C.c_uint32
        _lib.bsx_gate_num_constraints.restype = C.c_uint32
    return _lib


def _ptr(a) -> C.c_void_p:
    """numpy array -> host pointer; int -> raw (device) pointer; None -> NULL."""
    if a is None:
        return C.c_void_p(0)
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return a.ctypes.data_as(C.c_void_p)


def _in(a, dtype=np.uint8) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=dtype)


def pow2_ceil(n: int) -> int:
    p = 1
    while p < n:
        p *= 2
    return p


class Context:
    """One per GPU and per calling thread (a ctx is not thread-safe, like the reference's witness loop)."""

    def __init__(self, device: int = 0):
        self._lib = load()
        h = C.c_void_p()
        rc = self._lib.bsx_init(int(device), C.byref(h))
        if rc != 0:
            raise BsxError(f"bsx_init(device={device}) failed with status {rc} (no CUDA device? there is no CPU fallback)")
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.bsx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- plumbing --
    def _call(self, name: str, *args):
        rc = getattr(self._lib, name)(self._h, *args)
        if rc != 0:
            raise BsxError(f"{name} failed ({rc}): {self._lib.bsx_last_error(self._h).decode()}")

    @property
    def launch_count(self) -> int:
        return int(self._lib.bsx_launch_count(self._h))

    def sync(self):
        self._call("bsx_sync")

    def set_tunable(self, name: str, value: int):
        """Measurement knob of this ctx (include/bsx.h: bsx_set_tunable), e.g. ("ED_OCC", 8)."""
        self._call("bsx_set_tunable", C.c_char_p(name.encode()), C.c_int(int(value)))

    def get_tunable(self, name: str) -> int:
        v = C.c_int(0)
        self._call("bsx_get_tunable", C.c_char_p(name.encode()), C.byref(v))
        return int(v.value)

    # -- K1 --
    def sha256_batch(self, msgs, offsets) -> np.ndarray:
        offsets = _in(offsets, np.uint32)
        n = len(offsets) - 1
        out = np.zeros((n, 32), np.uint8)
        self._call("bsx_sha256_batch", _ptr(_in(msgs)), _ptr(offsets), C.c_uint32(n), _ptr(out))
        return out

    def sha512_batch(self, msgs, offsets) -> np.ndarray:
        offsets = _in(offsets, np.uint32)
        n = len(offsets) - 1
        out = np.zeros((n, 64), np.uint8)
        self._call("bsx_sha512_batch", _ptr(_in(msgs)), _ptr(offsets), C.c_uint32(n), _ptr(out))
        return out

    # -- K3 --
    def tm_merkle_proofs(self, leaves, leaf_len: int, aunts, depth: int, path_bits, hashed_leaf: bool = False):
        path_bits = _in(path_bits, np.uint32)
        n = len(path_bits)
        nd = 2 * depth + (0 if hashed_leaf else 1)
        dig = np.zeros((n, nd, 32), np.uint8)
        roots = np.zeros((n, 32), np.uint8)
        self._call("bsx_tm_merkle_proofs", _ptr(_in(leaves)), C.c_uint32(leaf_len), _ptr(_in(aunts)), C.c_uint32(depth),
                   _ptr(path_bits), C.c_uint32(n), C.c_int(int(hashed_leaf)), _ptr(dig), _ptr(roots))
        return dig, roots

    # -- K2 --
    def tm_merkle_tree(self, leaf_digests, N: int, nb_enabled):
        nb_enabled = _in(nb_enabled, np.uint64)
        t = len(nb_enabled)
        P = pow2_ceil(N)
        inner = np.zeros((t, P - 1, 32), np.uint8)
        roots = np.zeros((t, 32), np.uint8)
        self._call("bsx_tm_merkle_tree", _ptr(_in(leaf_digests)), C.c_uint32(N), C.c_uint32(t), _ptr(nb_enabled), _ptr(inner),
                   _ptr(roots))
        return inner, roots

    def data_commitment_batch(self, data_hashes, N: int, start_blocks, end_blocks):
        start_blocks, end_blocks = _in(start_blocks, np.uint64), _in(end_blocks, np.uint64)
        t = len(start_blocks)
        P = pow2_ceil(N)
        dig = np.zeros((t, N + P - 1, 32), np.uint8)
        roots = np.zeros((t, 32), np.uint8)
        fail = np.zeros(t, np.uint32)
        self._call("bsx_data_commitment_batch", _ptr(_in(data_hashes)), C.c_uint32(N), C.c_uint32(t), _ptr(start_blocks),
                   _ptr(end_blocks), _ptr(dig), _ptr(roots), _ptr(fail))
        return dig, roots, fail

    def attestation_proofs(self, data_hashes, N: int, start_blocks, end_blocks, q_tree, q_height):
        """Inclusion proofs of data-root tuples (BlobstreamX.verifyAttestation): per query (tree, height) ->
        dict(side_nodes [nq, max_depth, 32], depth, key, num_leaves, roots [t, 32])."""
        start_blocks, end_blocks = _in(start_blocks, np.uint64), _in(end_blocks, np.uint64)
        q_tree, q_height = _in(q_tree, np.uint32), _in(q_height, np.uint64)
        t, nq = len(start_blocks), len(q_tree)
        md = int(self._lib.bsx_attestation_max_depth(C.c_uint32(N)))
        out = dict(side_nodes=np.zeros((nq, max(md, 1), 32), np.uint8), depth=np.zeros(nq, np.uint32), key=np.zeros(nq, np.uint32),
                   num_leaves=np.zeros(nq, np.uint32), roots=np.zeros((t, 32), np.uint8))
        self._call("bsx_attestation_proofs", _ptr(_in(data_hashes)), C.c_uint32(N), C.c_uint32(t), _ptr(start_blocks), _ptr(end_blocks),
                   C.c_uint32(nq), _ptr(q_tree), _ptr(q_height), _ptr(out["side_nodes"]), _ptr(out["depth"]), _ptr(out["key"]),
                   _ptr(out["num_leaves"]), _ptr(out["roots"]))
        return out

    # -- map circuit --
    def prove_subchain_batch(self, B: int, dh_leaf, dh_aunts, lb_leaf, lb_aunts, start_headers, end_headers, batch_start,
                             batch_end, global_end, global_end_header):
        batch_start = _in(batch_start, np.uint64)
        n = len(batch_start)
        dig = np.zeros((n, 20 * B - 1, 32), np.uint8)
        sub = np.zeros((n, SUBCHAIN_BYTES), np.uint8)
        self._call("bsx_prove_subchain_batch", C.c_uint32(B), C.c_uint32(n), _ptr(_in(dh_leaf)), _ptr(_in(dh_aunts)),
                   _ptr(_in(lb_leaf)), _ptr(_in(lb_aunts)), _ptr(_in(start_headers)), _ptr(_in(end_headers)),
                   _ptr(batch_start), _ptr(_in(batch_end, np.uint64)), _ptr(_in(global_end, np.uint64)),
                   _ptr(_in(global_end_header)), _ptr(dig), _ptr(sub))
        return dig, sub

    def prove_data_commitment(self, n_ranges: int, n_jobs: int, B: int, dh_leaf, dh_aunts, lb_leaf, lb_aunts, start_headers,
                              end_headers, start_blocks, start_header, end_blocks, end_header):
        R = n_ranges
        out = dict(map_digests=np.zeros((R, n_jobs, 20 * B - 1, 32), np.uint8),
                   map_subchains=np.zeros((R, n_jobs, SUBCHAIN_BYTES), np.uint8),
                   reduce_digests=np.zeros((R, n_jobs - 1, 32), np.uint8),
                   reduce_nodes=np.zeros((R, n_jobs - 1, SUBCHAIN_BYTES), np.uint8),
                   data_commitments=np.zeros((R, 32), np.uint8), fail=np.zeros(R, np.uint32))
        self._call("bsx_prove_data_commitment", C.c_uint32(R), C.c_uint32(n_jobs), C.c_uint32(B), _ptr(_in(dh_leaf)),
                   _ptr(_in(dh_aunts)), _ptr(_in(lb_leaf)), _ptr(_in(lb_aunts)), _ptr(_in(start_headers)),
                   _ptr(_in(end_headers)), _ptr(_in(start_blocks, np.uint64)), _ptr(_in(start_header)),
                   _ptr(_in(end_blocks, np.uint64)), _ptr(_in(end_header)), _ptr(out["map_digests"]),
                   _ptr(out["map_subchains"]), _ptr(out["reduce_digests"]), _ptr(out["reduce_nodes"]),
                   _ptr(out["data_commitments"]), _ptr(out["fail"]))
        return out

    # -- K4+K5 --
    def ed25519_batch(self, pks, sigs, msgs, msg_lens=None, active=None) -> np.ndarray:
        """n signatures -> [n, 576] witness records (layout: include/bsx.h).  msgs is [n, stride]."""
        pks = _in(pks).reshape(-1, 32)
        n = pks.shape[0]
        msgs = _in(msgs).reshape(n, -1) if n else _in(msgs)
        stride = msgs.shape[1] if n else 0
        out = np.zeros((n, SIG_OUT_BYTES), np.uint8)
        lens = None if msg_lens is None else _in(msg_lens, np.uint32)
        act = None if active is None else _in(active)
        self._call("bsx_ed25519_batch", C.c_uint32(n), _ptr(pks), _ptr(_in(sigs)), _ptr(msgs), C.c_uint32(stride), _ptr(lens),
                   _ptr(act), _ptr(out))
        return out

    # -- K6/K7/K8 --
    def gate_num_wires(self, gate: int, p0: int, p1: int = 0) -> int:
        return int(self._lib.bsx_gate_num_wires(C.c_uint32(gate), C.c_uint32(p0), C.c_uint32(p1)))

    def gate_num_constraints(self, gate: int, p0: int, p1: int = 0) -> int:
        return int(self._lib.bsx_gate_num_constraints(C.c_uint32(gate), C.c_uint32(p0), C.c_uint32(p1)))

    def gl_gate_eval(self, gate: int, p0: int, p1: int, wires) -> np.ndarray:
        """wires [n_wires, rows] u64 wire-major -> constraints [n_constraints, rows]."""
        wires = _in(wires, np.uint64)
        assert wires.shape[0] == self.gate_num_wires(gate, p0, p1)
        rows = wires.shape[1]
        out = np.zeros((self.gate_num_constraints(gate, p0, p1), rows), np.uint64)
        self._call("bsx_gl_gate_eval", C.c_uint32(gate), C.c_uint32(p0), C.c_uint32(p1), _ptr(wires), C.c_uint32(rows), _ptr(out))
        return out

    def gl_gate_witness(self, gate: int, p0: int, p1: int, wires) -> np.ndarray:
        wires = _in(wires, np.uint64).copy()
        self._call("bsx_gl_gate_witness", C.c_uint32(gate), C.c_uint32(p0), C.c_uint32(p1), _ptr(wires), C.c_uint32(wires.shape[1]))
        return wires

    def gl_poseidon_batch(self, inputs, offsets) -> np.ndarray:
        offsets = _in(offsets, np.uint32)
        n = len(offsets) - 1
        out = np.zeros((n, 4), np.uint64)
        self._call("bsx_gl_poseidon_batch", _ptr(_in(inputs, np.uint64)), _ptr(offsets), C.c_uint32(n), _ptr(out))
        return out

    # -- input shaping on the device --
    def header_trees(self, records, levels: bool = False):
        """Header hashes (and optionally all 27 tree digests) of n header records (inputs.pack_header_record)."""
        rec = _in(records).reshape(-1, 512)
        n = len(rec)
        roots = np.zeros((n, 32), np.uint8)
        lv = np.zeros((n, 27, 32), np.uint8) if levels else None
        self._call("bsx_header_trees", _ptr(rec), C.c_uint32(n), _ptr(roots), _ptr(lv))
        return (roots, lv) if levels else roots

    def header_range_inputs(self, records, start_blocks, end_blocks, n_jobs: int, batch_size: int, latest_blocks=None) -> dict:
        """Map-circuit inputs of n ranges from their header records [n, n_jobs*B+1, 512] -> the input arrays of
        header_range / prove_data_commitment (dh_leaf, dh_aunts, lb_leaf, lb_aunts, start_headers, end_headers,
        start_header, end_header) + fail[n].  latest_blocks: last fetchable block per range (None: the range's end)."""
        sb, eb = _in(start_blocks, np.uint64).reshape(-1), _in(end_blocks, np.uint64).reshape(-1)
        lt = None if latest_blocks is None else _in(latest_blocks, np.uint64).reshape(-1)
        n, J, B = len(sb), n_jobs, batch_size
        rec = _in(records).reshape(n, J * B + 1, 512)
        out = dict(dh_leaf=np.zeros((n, J * B, 34), np.uint8), dh_aunts=np.zeros((n, J * B, 128), np.uint8),
                   lb_leaf=np.zeros((n, J * B, 72), np.uint8), lb_aunts=np.zeros((n, J * B, 128), np.uint8),
                   start_headers=np.zeros((n, J, 32), np.uint8), end_headers=np.zeros((n, J, 32), np.uint8),
                   start_header=np.zeros((n, 32), np.uint8), end_header=np.zeros((n, 32), np.uint8), fail=np.zeros(n, np.uint32))
        self._call("bsx_header_range_inputs", C.c_uint32(n), C.c_uint32(J), C.c_uint32(B), _ptr(rec), _ptr(sb), _ptr(eb),
                   _ptr(lt) if lt is not None else C.c_void_p(0), *[_ptr(out[k]) for k in ("dh_leaf", "dh_aunts", "lb_leaf", "lb_aunts", "start_headers", "end_headers",
                                            "start_header", "end_header", "fail")])
        return out

    def encode_headers(self, fields) -> np.ndarray:
        """bsx_header_fields records (inputs.pack_header_fields) -> header records [n, 512]."""
        f = np.ascontiguousarray(fields)
        assert f.dtype.itemsize == 464
        out = np.zeros((f.size, 512), np.uint8)
        self._call("bsx_encode_headers", C.c_uint32(f.size), _ptr(f), _ptr(out))
        return out

    def validator_records(self, commits, sigs, n_validators: int, records: bool = True, hash_fields: bool = False) -> dict:
        """commits [n] (bsx_commit_in) + sigs [n, N] (bsx_commit_sig_in), from inputs.pack_commit ->
        validators [n, N, 240] and/or pubkeys [n, N, 32], powers [n, N], byte_lengths [n, N]; fail [n]."""
        cm, sg = np.ascontiguousarray(commits).reshape(-1), np.ascontiguousarray(sigs)
        n, N = cm.size, n_validators
        assert cm.dtype.itemsize == 152 and sg.dtype.itemsize == 160 and sg.size == n * N
        out = dict(fail=np.zeros(n, np.uint32))
        if records:
            out["validators"] = np.zeros((n, N, 240), np.uint8)
        if hash_fields:
            out.update(pubkeys=np.zeros((n, N, 32), np.uint8), powers=np.zeros((n, N), np.uint64), byte_lengths=np.zeros((n, N), np.uint32))
        self._call("bsx_validator_records", C.c_uint32(n), C.c_uint32(N), _ptr(cm), _ptr(sg), _ptr(out.get("validators")),
                   _ptr(out.get("pubkeys")), _ptr(out.get("powers")), _ptr(out.get("byte_lengths")), _ptr(out["fail"]))
        return out

    def present_on_trusted(self, target_sigs, n_target, trusted_sigs, n_trusted, validators):
        """Sets present_on_trusted_header in validators [n, N, 240] (in place) -> fail [n]."""
        val = validators
        assert val.dtype == np.uint8 and val.flags.c_contiguous and val.ndim == 3
        n, N = val.shape[0], val.shape[1]
        tg, tr = np.ascontiguousarray(target_sigs), np.ascontiguousarray(trusted_sigs)
        assert tg.size == n * N and tr.size == n * N and tg.dtype.itemsize == 160
        nt, ns = _in(n_target, np.uint32).reshape(-1), _in(n_trusted, np.uint32).reshape(-1)
        fail = np.zeros(n, np.uint32)
        self._call("bsx_present_on_trusted", C.c_uint32(n), C.c_uint32(N), _ptr(tg), _ptr(nt), _ptr(tr), _ptr(ns), _ptr(val), _ptr(fail))
        return fail

    # -- witness data formats --
    def hash_input_data(self, bufs, buf_offsets, lens, kinds, sha512: bool = False):
        """HashInputData of one SHA accelerator -> dict(padded_chunks [chunks,16], end_bits, digest_bits, digest_indices)."""
        buf_offsets, lens, kinds = _in(buf_offsets, np.uint32), _in(lens, np.uint32), _in(kinds, np.uint8)
        n = len(kinds)
        total = C.c_uint32(0)
        a = (C.c_int(int(sha512)), C.c_uint32(n), _ptr(_in(bufs)), _ptr(buf_offsets), _ptr(lens), _ptr(kinds))
        self._call("bsx_hash_input_data", *a, _ptr(None), _ptr(None), _ptr(None), _ptr(None), C.byref(total))
        t = total.value
        pc = np.zeros((t, 16), np.uint64 if sha512 else np.uint32)
        eb, db, di = np.zeros(t, np.uint8), np.zeros(t, np.uint8), np.zeros(n, np.uint32)
        if n:
            self._call("bsx_hash_input_data", *a, _ptr(pc), _ptr(eb), _ptr(db), _ptr(di), C.byref(total))
        return dict(padded_chunks=pc, end_bits=eb, digest_bits=db, digest_indices=di)

    def witness_pack_bytes(self, data) -> np.ndarray:
        data = _in(data).reshape(-1)
        out = np.zeros((len(data), 8), np.uint64)
        self._call("bsx_witness_pack_bytes", _ptr(data), C.c_size_t(len(data)), _ptr(out))
        return out

    def witness_unpack_bytes(self, elements):
        el = _in(elements, np.uint64).reshape(-1, 8)
        out = np.zeros(len(el), np.uint8)
        bad = C.c_uint32(0)
        self._call("bsx_witness_unpack_bytes", _ptr(el), C.c_size_t(len(el)), _ptr(out), C.byref(bad))
        return out, bool(bad.value)

    # -- verify_header / verify_skip / next_header --
    def _verify(self, mode: int, items, N: int):
        """items: list of dicts as produced by blobstreamx_b200.inputs.get_skip_inputs / get_step_inputs /
        _header_common (one per instance).  Returns dict(sha256_digests [n,D,32], ed [n,N,576], fail [n], ...)."""
        n = len(items)
        key = {0: None, 1: "target", 2: "next"}[mode]
        hdr = pack_header_in([it[key] if key else it for it in items])
        vals = np.stack([_in((it[key] if key else it)["validators"]).reshape(N, VAL_IN_BYTES) for it in items])
        D = int(self._lib.bsx_verify_digest_count(C.c_int(mode), C.c_uint32(N)))
        dig = np.zeros((n, D, 32), np.uint8)
        ed = np.zeros((n, N, SIG_OUT_BYTES), np.uint8)
        fail = np.zeros(n, np.uint32)
        out = dict(sha256_digests=dig, ed=ed, fail=fail)
        if mode == 0:
            self._call("bsx_verify_header", C.c_uint32(n), C.c_uint32(N), _ptr(hdr), _ptr(vals), _ptr(dig), _ptr(ed), _ptr(fail))
        elif mode == 1:
            skip = pack_skip_in(items)
            tpk = np.stack([_in(it["trusted_pubkeys"]).reshape(N, 32) for it in items])
            tpw = np.stack([_in(it["trusted_powers"], np.uint64) for it in items])
            tbl = np.stack([_in(it["trusted_byte_lengths"], np.uint32) for it in items])
            self._call("bsx_verify_skip", C.c_uint32(n), C.c_uint32(N), _ptr(hdr), _ptr(vals), _ptr(skip), _ptr(tpk), _ptr(tpw),
                       _ptr(tbl), _ptr(dig), _ptr(ed), _ptr(fail))
        else:
            step = pack_step_in(items)
            dc = np.zeros((n, 32), np.uint8)
            self._call("bsx_next_header", C.c_uint32(n), C.c_uint32(N), _ptr(hdr), _ptr(vals), _ptr(step), _ptr(dig), _ptr(ed),
                       _ptr(dc), _ptr(fail))
            out["data_commitments"] = dc
        return out

    def verify_header(self, items, N: int = 100):
        return self._verify(0, items, N)

    def verify_skip(self, items, N: int = 100):
        return self._verify(1, items, N)

    def next_header(self, items, N: int = 100):
        return self._verify(2, items, N)

    # -- header_range = skip + prove_data_commitment --
    def header_range(self, skip_items, m, N: int = 100):
        """One or more header ranges.  skip_items: list of get_skip_inputs dicts; m: dict of the flat map arrays
        (as bench.tile_ranges / inputs.HeaderRangeMapInputs fields) incl. n_jobs, batch."""
        n, J, B = len(skip_items), int(m["n_jobs"]), int(m["batch"])
        keep = dict(
            hdr=pack_header_in([k["target"] for k in skip_items]),
            validators=np.stack([_in(k["target"]["validators"]).reshape(N, VAL_IN_BYTES) for k in skip_items]),
            skip=pack_skip_in(skip_items),
            trusted_pubkeys=np.stack([_in(k["trusted_pubkeys"]).reshape(N, 32) for k in skip_items]),
            trusted_powers=np.stack([_in(k["trusted_powers"], np.uint64) for k in skip_items]),
            trusted_byte_lengths=np.stack([_in(k["trusted_byte_lengths"], np.uint32) for k in skip_items]))
        D = int(self._lib.bsx_verify_digest_count(C.c_int(1), C.c_uint32(N)))
        so = dict(digests=np.zeros((n, D, 32), np.uint8), ed_out=np.zeros((n, N, SIG_OUT_BYTES), np.uint8), fail=np.zeros(n, np.uint32))
        sb = fill_struct(SkipBatch(), **keep, **so)
        mi = {f: _in(m[f]) for f in ("dh_leaf", "dh_aunts", "lb_leaf", "lb_aunts", "start_headers", "end_headers", "start_header", "end_header")}
        mi["start_blocks"], mi["end_blocks"] = _in(m["start_blocks"], np.uint64), _in(m["end_blocks"], np.uint64)
        mo = dict(map_digests=np.zeros((n, J, 20 * B - 1, 32), np.uint8), map_subchains=np.zeros((n, J, SUBCHAIN_BYTES), np.uint8),
                  reduce_digests=np.zeros((n, max(J - 1, 1), 32), np.uint8), reduce_nodes=np.zeros((n, max(J - 1, 1), SUBCHAIN_BYTES), np.uint8),
                  data_commitments=np.zeros((n, 32), np.uint8), fail=np.zeros(n, np.uint32))
        rb = fill_struct(RangeBatch(), **mi, **mo)
        self._call("bsx_header_range", C.c_uint32(n), C.c_uint32(N), C.c_uint32(J), C.c_uint32(B), C.byref(sb), C.byref(rb))
        return dict(skip=dict(sha256_digests=so["digests"], ed=so["ed_out"], fail=so["fail"]), **mo)

    # -- raw access for device-pointer entry points (bench / multi-GPU) --
    def call_dev(self, name: str, stream: int, *args):
        """Call a `*_dev` entry point; integer args that are pointers must be wrapped with ptr()."""
        self._call(name, C.c_void_p(int(stream)), *args)


class SkipBatch(C.Structure):
    """bsx_skip_batch"""
    _fields_ = [(k, C.c_void_p) for k in ("hdr", "validators", "skip", "trusted_pubkeys", "trusted_powers",
                                          "trusted_byte_lengths", "digests", "ed_out", "fail")]


class RangeBatch(C.Structure):
    """bsx_range_batch"""
    _fields_ = [(k, C.c_void_p) for k in ("dh_leaf", "dh_aunts", "lb_leaf", "lb_aunts", "start_headers", "end_headers",
                                          "start_blocks", "end_blocks", "start_header", "end_header", "map_digests",
                                          "map_subchains", "reduce_digests", "reduce_nodes", "data_commitments", "fail")]


class ShardIn(C.Structure):
    """bsx_shard_in"""
    _fields_ = [(k, C.c_void_p) for k in ("dh_leaf", "dh_aunts", "lb_leaf", "lb_aunts", "start_headers", "end_headers",
                                          "batch_start", "batch_end", "global_end", "global_end_header", "start_blocks",
                                          "end_blocks", "start_header", "end_header")]


class ShardOut(C.Structure):
    """bsx_shard_out"""
    _fields_ = [(k, C.c_void_p) for k in ("map_digests", "map_subchains", "reduce_digests", "reduce_nodes", "data_commitments", "fail")]


IPC_HANDLE_BYTES = 64


class Shard:
    """bsx_shard: one rank of the sharded header_range map/reduce (include/bsx.h).  Setup: exchange `ipc_handle()` between
    processes and `open_peer`, or `exchange_buffer()` pointers inside one process and `set_peer`; then `step_dev` per step."""

    def __init__(self, ctx: "Context", rank: int, world: int, n_ranges: int, n_jobs: int, batch: int, exchange_buf: int = 0):
        self.ctx, self.rank, self.world = ctx, rank, world
        lib_ = ctx._lib
        lib_.bsx_shard_exchange_bytes.restype = C.c_size_t
        h = C.c_void_p()
        ctx._call("bsx_shard_create", C.c_uint32(rank), C.c_uint32(world), C.c_uint32(n_ranges), C.c_uint32(n_jobs), C.c_uint32(batch),
                  C.c_void_p(int(exchange_buf)), C.byref(h))
        self._h = h

    @staticmethod
    def exchange_bytes(world: int, n_ranges: int, n_jobs: int) -> int:
        lib_ = load()
        lib_.bsx_shard_exchange_bytes.restype = C.c_size_t
        return int(lib_.bsx_shard_exchange_bytes(C.c_uint32(world), C.c_uint32(n_ranges), C.c_uint32(n_jobs)))

    def _call(self, name, *args):
        rc = getattr(self.ctx._lib, name)(self._h, *args)
        if rc != 0:
            raise BsxError(f"{name} failed ({rc}): {self.ctx._lib.bsx_last_error(self.ctx._h).decode()}")

    def exchange_buffer(self):
        p, n = C.c_void_p(), C.c_size_t()
        self._call("bsx_shard_exchange_buffer", C.byref(p), C.byref(n))
        return int(p.value), int(n.value)

    def ipc_handle(self) -> bytes:
        b = (C.c_uint8 * IPC_HANDLE_BYTES)()
        self._call("bsx_shard_ipc_handle", b)
        return bytes(b)

    def open_peer(self, peer: int, handle: bytes):
        self._call("bsx_shard_open_peer", C.c_uint32(peer), (C.c_uint8 * IPC_HANDLE_BYTES).from_buffer_copy(handle))

    def set_peer(self, peer: int, dev_ptr: int):
        self._call("bsx_shard_set_peer", C.c_uint32(peer), C.c_void_p(int(dev_ptr)))

    def step_dev(self, stream: int, sin: "ShardIn", sout: "ShardOut"):
        self._call("bsx_shard_step_dev", C.c_void_p(int(stream)), C.byref(sin), C.byref(sout))

    def close(self):
        if getattr(self, "_h", None) and getattr(self.ctx, "_h", None):
            self.ctx._lib.bsx_shard_destroy(self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _addr(a) -> int:
    if a is None:
        return 0
    if isinstance(a, (int, np.integer)):
        return int(a)
    return a.ctypes.data


def fill_struct(st, **ptrs):
    """numpy arrays / raw device addresses -> the pointer fields of a SkipBatch / RangeBatch."""
    for k, v in ptrs.items():
        setattr(st, k, _addr(v))
    return st


# struct layouts of include/bsx.h
HEADER_IN = np.dtype([("header", "u1", 32), ("height", "<u8"), ("round", "<u8"), ("nb_enabled", "<u8"),
                      ("chain_id_enc", "u1", 64), ("chain_id_enc_len", "<u4"), ("height_enc_len", "<u4"),
                      ("chain_id_aunts", "u1", 128), ("height_aunts", "u1", 128), ("validators_hash_proof", "u1", 168),
                      ("expected_chain_id", "u1", 56), ("expected_chain_id_len", "<u4"), ("_pad", "<u4")])
SKIP_IN = np.dtype([("trusted_block", "<u8"), ("trusted_nb_enabled", "<u8"), ("skip_max", "<u4"), ("_pad", "<u4"),
                    ("trusted_header", "u1", 32), ("trusted_validators_hash_proof", "u1", 168)])
STEP_IN = np.dtype([("prev_block", "<u8"), ("prev_header", "u1", 32), ("last_block_id_proof", "u1", 200),
                    ("prev_next_validators_proof", "u1", 168), ("data_hash_proof", "u1", 168)])
assert HEADER_IN.itemsize == 616 and SKIP_IN.itemsize == 224 and STEP_IN.itemsize == 576


def _put(dst, src):
    src = np.asarray(src, np.uint8).reshape(-1)
    dst[: len(src)] = src


def pack_header_in(hs) -> np.ndarray:
    a = np.zeros(len(hs), HEADER_IN)
    for i, h in enumerate(hs):
        r = a[i]
        _put(r["header"], h["header"])
        r["height"], r["round"], r["nb_enabled"] = h["height"], h["round"], h["nb_enabled"]
        _put(r["chain_id_enc"], h["chain_id_enc"])
        r["chain_id_enc_len"], r["height_enc_len"] = h["chain_id_enc_len"], h["height_enc_len"]
        _put(r["chain_id_aunts"], h["chain_id_aunts"])
        _put(r["height_aunts"], h["height_aunts"])
        _put(r["validators_hash_proof"], h["validators_hash_proof"])
        _put(r["expected_chain_id"], h["expected_chain_id"])
        r["expected_chain_id_len"] = len(h["expected_chain_id"])
    return a


def pack_skip_in(ks) -> np.ndarray:
    a = np.zeros(len(ks), SKIP_IN)
    for i, k in enumerate(ks):
        r = a[i]
        r["trusted_block"], r["trusted_nb_enabled"], r["skip_max"] = k["trusted_block"], k["trusted_nb_enabled"], k["skip_max"]
        _put(r["trusted_header"], k["trusted_header"])
        _put(r["trusted_validators_hash_proof"], k["trusted_validators_hash_proof"])
    return a


def pack_step_in(ks) -> np.ndarray:
    a = np.zeros(len(ks), STEP_IN)
    for i, k in enumerate(ks):
        r = a[i]
        r["prev_block"] = k["prev_block"]
        _put(r["prev_header"], k["prev_header"])
        _put(r["last_block_id_proof"], k["last_block_id_proof"])
        _put(r["prev_next_validators_proof"], k["prev_next_validators_proof"])
        _put(r["data_hash_proof"], k["data_hash_proof"])
    return a


def ptr(x) -> C.c_void_p:
    return _ptr(x)


def u32(x) -> C.c_uint32:
    return C.c_uint32(int(x))


def u64(x) -> C.c_uint64:
    return C.c_uint64(int(x))

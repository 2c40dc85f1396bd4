"""Host-side mirror of the prover inner loops (include/bsx.h "Prover inner loops over Goldilocks", csrc/k_plonk.cu):
the shape of plonky2's `PolynomialBatch::from_values` / `compute_quotient_polys` / FRI fold call sequence
(un-vendored; call site PX/backend/circuit/build.rs:69-75) over device-resident torch tensors.  torch only owns the
memory and the stream; every operation is a libbsx kernel.  uint64 field elements travel as torch.int64 bit patterns.

Layouts (as in bsx.h): polynomials are rows of a [n_polys, n] tensor; transforms produce BIT-REVERSED index order;
an extension of rate 2^r is [n_polys, n * 2^r] with position i = value at shift * w_N^bitrev(i)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import lib as L


def bitrev_indices(log_n: int) -> np.ndarray:
    """perm[i] = bitrev(i): natural[perm] = bit-reversed buffer read as natural order"""
    i = np.arange(1 << log_n, dtype=np.uint32)
    r = np.zeros_like(i)
    for b in range(log_n):
        r |= ((i >> b) & 1) << (log_n - 1 - b)
    return r.astype(np.int64)


class Prover:
    def __init__(self, ctx: L.Context, device: torch.device):
        self.ctx, self.dev = ctx, device
        lb = ctx._lib
        lb.bsx_gl_root_of_unity.restype = C.c_uint64
        lb.bsx_gl_coset_shift.restype = C.c_uint64
        lb.bsx_gl_merkle_digest_words.restype = C.c_size_t

    @property
    def stream(self) -> int:
        return torch.cuda.current_stream().cuda_stream

    def root_of_unity(self, log_n: int) -> int:
        return int(self.ctx._lib.bsx_gl_root_of_unity(C.c_uint32(log_n)))

    def coset_shift(self) -> int:
        return int(self.ctx._lib.bsx_gl_coset_shift())

    def ntt(self, x: torch.Tensor, inverse: bool = False, natural_out: bool = False) -> torch.Tensor:
        """[n_polys, n] -> transform of every row (natural in; bit-reversed out unless natural_out)"""
        n_polys, n = x.shape
        log_n = n.bit_length() - 1
        out = torch.empty_like(x)
        scratch = torch.empty_like(x) if natural_out else None
        self.ctx.call_dev("bsx_gl_ntt_dev", self.stream, L.ptr(x.data_ptr()), L.ptr(out.data_ptr()), L.u32(log_n), L.u32(n_polys),
                          C.c_size_t(x.stride(0)), C.c_size_t(out.stride(0)), C.c_int(int(inverse)), C.c_int(int(natural_out)),
                          L.ptr(scratch.data_ptr() if natural_out else 0))
        return out

    def lde(self, coeffs: torch.Tensor, rate_bits: int, shift: int = 0, out: torch.Tensor = None) -> torch.Tensor:
        n_polys, n = coeffs.shape
        log_n = n.bit_length() - 1
        if out is None:
            out = torch.empty((n_polys, n << rate_bits), dtype=torch.int64, device=self.dev)
        self.ctx.call_dev("bsx_gl_lde_dev", self.stream, L.ptr(coeffs.data_ptr()), L.ptr(out.data_ptr()), L.u32(log_n), L.u32(rate_bits),
                          L.u32(n_polys), C.c_size_t(coeffs.stride(0)), C.c_size_t(out.stride(0)), L.u64(shift))
        return out

    def merkle_caps(self, data: torch.Tensor, cap_height: int, out: torch.Tensor = None):
        """[width, n_leaves] poly-major -> (all digests [words/4, 4], cap [2^cap_height, 4])"""
        width, n_leaves = data.shape
        words = int(self.ctx._lib.bsx_gl_merkle_digest_words(C.c_uint32(n_leaves), C.c_uint32(cap_height)))
        if out is None:
            out = torch.empty(words, dtype=torch.int64, device=self.dev)
        self.ctx.call_dev("bsx_gl_merkle_caps_dev", self.stream, L.ptr(data.data_ptr()), C.c_size_t(data.stride(0)), L.u32(width),
                          L.u32(n_leaves), L.u32(cap_height), L.ptr(out.data_ptr()))
        d = out.view(-1, 4)
        return d, d[-(1 << cap_height):]

    def quotient_tables(self, alphas, n_constraints: int, log_n: int, rate_bits: int, shift: int = 0):
        al = np.ascontiguousarray(alphas, np.uint64)
        ap = torch.empty(len(al) * n_constraints, dtype=torch.int64, device=self.dev)
        zh = torch.empty(1 << rate_bits, dtype=torch.int64, device=self.dev)
        self.ctx.call_dev("bsx_gl_quotient_tables_dev", self.stream, L.ptr(al.ctypes.data), L.u32(len(al)), L.u32(n_constraints),
                          L.u32(log_n), L.u32(rate_bits), L.u64(shift), L.ptr(ap.data_ptr()), L.ptr(zh.data_ptr()))
        return ap, zh

    def gate_quotient(self, gate: int, p0: int, p1: int, lde: torch.Tensor, alpha_pows: torch.Tensor, n_alphas: int, zh_inv: torch.Tensor,
                      log_block: int, out: torch.Tensor = None) -> torch.Tensor:
        """lde: [n_wires, rows] (contiguous rows) -> [n_alphas, rows]"""
        rows = lde.shape[1]
        assert lde.stride(0) == rows
        if out is None:
            out = torch.empty((n_alphas, rows), dtype=torch.int64, device=self.dev)
        self.ctx.call_dev("bsx_gl_gate_quotient_dev", self.stream, L.u32(gate), L.u32(p0), L.u32(p1), L.ptr(lde.data_ptr()), L.u32(rows),
                          L.ptr(alpha_pows.data_ptr()), L.u32(n_alphas), L.ptr(zh_inv.data_ptr()), L.u32(log_block), L.ptr(out.data_ptr()))
        return out

    def fri_fold(self, pairs: torch.Tensor, arity_bits: int, beta) -> torch.Tensor:
        """[n, 2] extension elements -> [n >> arity_bits, 2]"""
        n = pairs.shape[0]
        out = torch.empty((n >> arity_bits, 2), dtype=torch.int64, device=self.dev)
        self.ctx.call_dev("bsx_gl_fri_fold_dev", self.stream, L.ptr(pairs.data_ptr()), L.u32(n), L.u32(arity_bits), L.u64(int(beta[0])),
                          L.u64(int(beta[1])), L.ptr(out.data_ptr()))
        return out

    def sha256_trace(self, padded_chunks: torch.Tensor, end_bits: torch.Tensor, digest_bits: torch.Tensor, log_rows: int,
                     out: torch.Tensor = None) -> torch.Tensor:
        """HashInputData of one SHA-256 accelerator (device tensors: [chunks, 16] int32 words, uint8 flags) -> the execution
        trace [BSX_SHA256_TRACE_COLS, 2^log_rows] (include/bsx.h; replaces the row-by-row fill of HashStark::prove,
        PX/frontend/hash/curta/stark.rs:107-133; layout our own, parity unpinned)."""
        n = padded_chunks.shape[0]
        if out is None:
            out = torch.empty((SHA256_TRACE_COLS, 1 << log_rows), dtype=torch.int64, device=self.dev)
        self.ctx.call_dev("bsx_sha256_trace_dev", self.stream, L.ptr(padded_chunks.data_ptr()), L.ptr(end_bits.data_ptr()),
                          L.ptr(digest_bits.data_ptr()), L.u32(n), L.u32(log_rows), L.ptr(out.data_ptr()))
        return out

    def sha256_trace_batch(self, padded_chunks: torch.Tensor, end_bits: torch.Tensor, digest_bits: torch.Tensor, log_rows: int,
                           out: torch.Tensor = None) -> torch.Tensor:
        """[circuits, chunks, 16] int32 words + [circuits, chunks] uint8 flags (same request schedule in every circuit) ->
        [circuits, BSX_SHA256_TRACE_COLS, 2^log_rows] in one launch."""
        nc, n = padded_chunks.shape[0], padded_chunks.shape[1]
        if out is None:
            out = torch.empty((nc, SHA256_TRACE_COLS, 1 << log_rows), dtype=torch.int64, device=self.dev)
        self.ctx.call_dev("bsx_sha256_trace_batch_dev", self.stream, L.ptr(padded_chunks.data_ptr()), L.ptr(end_bits.data_ptr()),
                          L.ptr(digest_bits.data_ptr()), L.u32(n), L.u32(nc), C.c_size_t(n), L.u32(log_rows), L.ptr(out.data_ptr()))
        return out

    def sha512_trace(self, padded_chunks: torch.Tensor, end_bits: torch.Tensor, digest_bits: torch.Tensor, log_rows: int) -> torch.Tensor:
        """As sha256_trace for the SHA-512 (EdDSA) accelerator: [chunks, 16] int64 words -> [BSX_SHA512_TRACE_COLS, 2^log_rows]."""
        out = torch.empty((SHA512_TRACE_COLS, 1 << log_rows), dtype=torch.int64, device=self.dev)
        self.ctx.call_dev("bsx_sha512_trace_dev", self.stream, L.ptr(padded_chunks.data_ptr()), L.ptr(end_bits.data_ptr()),
                          L.ptr(digest_bits.data_ptr()), L.u32(padded_chunks.shape[0]), L.u32(log_rows), L.ptr(out.data_ptr()))
        return out

    def ed25519_trace_scratch_bytes(self, n: int) -> int:
        lb = self.ctx._lib
        lb.bsx_ed25519_trace_scratch_bytes.restype = C.c_size_t
        return max(int(lb.bsx_ed25519_trace_scratch_bytes(C.c_uint32(n))), 256)

    def ed25519_trace_scratch(self, n: int) -> torch.Tensor:
        return torch.empty(self.ed25519_trace_scratch_bytes(n), dtype=torch.uint8, device=self.dev)

    def ed25519_trace(self, scalars: torch.Tensor, points: torch.Tensor, log_rows: int, out: torch.Tensor = None, results: bool = True):
        """n scalar multiplications k * P (device tensors: [n, 32] uint8 little-endian scalars, [n, 64] uint8 affine points,
        the s / G and h / A of the EdDSA schedule) -> (trace [BSX_ED25519_TRACE_COLS, 2^log_rows], [n, 64] uint8 k * P).
        Replaces the trace fill of Ed25519Stark::prove (PX/frontend/ecc/curve25519/curta/stark.rs:182-219); layout our own,
        parity unpinned (include/bsx.h)."""
        n = scalars.shape[0]
        if getattr(self, "_edt_scratch", None) is None or self._edt_scratch.numel() < self.ed25519_trace_scratch_bytes(n):
            self._edt_scratch = self.ed25519_trace_scratch(n)
        if out is None:
            out = torch.empty((ED25519_TRACE_COLS, 1 << log_rows), dtype=torch.int64, device=self.dev)
        res = torch.empty((n, 64), dtype=torch.uint8, device=self.dev) if results else None
        self.ctx.call_dev("bsx_ed25519_trace_dev", self.stream, L.ptr(scalars.data_ptr() if n else 0), L.ptr(points.data_ptr() if n else 0),
                          L.u32(n), L.u32(log_rows), L.ptr(self._edt_scratch.data_ptr()), L.ptr(res.data_ptr() if results and n else 0),
                          L.ptr(out.data_ptr()))
        return out, res

    def ed25519_trace_operands(self, sigs: torch.Tensor, ed_out: torch.Tensor, active: torch.Tensor = None):
        """[n, 64] uint8 signatures (+ [n] uint8 lane flags) + [n, 576] uint8 witness records (device; rows may be strided views,
        e.g. validators[:, 32:96]) -> ([2n, 32] scalars, [2n, 64] points): (s, G), (h, A); inactive lanes take the DUMMY s"""
        n = sigs.shape[0]
        scalars = torch.empty((2 * n, 32), dtype=torch.uint8, device=self.dev)
        points = torch.empty((2 * n, 64), dtype=torch.uint8, device=self.dev)
        self.ctx.call_dev("bsx_ed25519_trace_operands_dev", self.stream, L.u32(n), L.ptr(sigs.data_ptr()), L.u32(sigs.stride(0)),
                          L.ptr(active.data_ptr() if active is not None else 0), L.u32(active.stride(0) if active is not None else 0),
                          L.ptr(ed_out.data_ptr()), L.ptr(scalars.data_ptr()), L.ptr(points.data_ptr()))
        return scalars, points

    def ed25519_trace_points(self, scalars: torch.Tensor, points: torch.Tensor, scratch: torch.Tensor, stream: int = None):
        """first half of ed25519_trace (multiplication chains -> scratch), on `stream`"""
        self.ctx.call_dev("bsx_ed25519_trace_points_dev", self.stream if stream is None else stream, L.ptr(scalars.data_ptr()),
                          L.ptr(points.data_ptr()), L.u32(scalars.shape[0]), L.ptr(scratch.data_ptr()))

    def ed25519_trace_rows(self, scalars: torch.Tensor, points: torch.Tensor, scratch: torch.Tensor, log_rows: int, out: torch.Tensor,
                           stream: int = None):
        """second half of ed25519_trace (scratch -> trace), on `stream`"""
        self.ctx.call_dev("bsx_ed25519_trace_rows_dev", self.stream if stream is None else stream, L.ptr(scalars.data_ptr()),
                          L.ptr(points.data_ptr()), L.u32(scalars.shape[0]), L.u32(log_rows), L.ptr(scratch.data_ptr()), L.ptr(0), L.ptr(out.data_ptr()))


SHA256_TRACE_COLS = 176
ED25519_TRACE_COLS = 1540
SHA512_TRACE_COLS = 338

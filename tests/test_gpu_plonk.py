"""GPU parity for the prover inner loops (k_plonk.cu, gl_gate_quotient_kernel) against the CPU restatement
(oracle/plonk.c, natural index order): transforms of every pass plan (1, 2 and 3 passes over HBM), coset extension,
Poseidon Merkle caps, quotient over the extension, FRI fold -- and the chain trace -> coefficients -> extension ->
{Merkle cap, quotient} run end to end on the device with a VALID trace, whose quotient must be a polynomial of degree < 3n."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 2**64 - 2**32 + 1


@pytest.fixture(scope="module")
def pv():
    import torch
    from blobstreamx_b200 import lib
    from blobstreamx_b200.plonk import Prover
    ctx = lib.Context(0)
    yield Prover(ctx, torch.device("cuda", 0))
    ctx.close()


def _dev(pv, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).to(pv.dev)


def _host(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.mark.parametrize("log_n,n_polys", [(1, 3), (5, 2), (10, 3), (11, 5), (12, 2), (16, 3), (20, 2), (21, 1)])
def test_ntt_forward_inverse(pv, log_n, n_polys):
    from blobstreamx_b200.plonk import bitrev_indices
    from oracle import cbind as orc
    rng = np.random.default_rng(log_n)
    x = rng.integers(0, 2**64, (n_polys, 1 << log_n), dtype=np.uint64)          # any 64-bit representative
    xc = (x.astype(object) % P).astype(np.uint64)
    got = _host(pv.ntt(_dev(pv, x)))
    assert (got < P).all()
    perm = bitrev_indices(log_n)
    if log_n <= 16:
        want = orc.gl_ntt(xc)
        assert (got == want[:, perm]).all()
        assert (_host(pv.ntt(_dev(pv, x), natural_out=True)) == want).all()
    # inverse of the forward transform (fed back in natural order) is the identity, at every size
    nat = np.empty_like(got)
    nat[:, perm] = got
    back = _host(pv.ntt(_dev(pv, nat), inverse=True, natural_out=True))
    assert (back == xc).all()


@pytest.mark.parametrize("log_n,rate_bits,n_polys", [(3, 3, 2), (10, 3, 3), (12, 1, 2), (13, 3, 4), (16, 3, 2)])
def test_lde_matches_oracle(pv, log_n, rate_bits, n_polys):
    from blobstreamx_b200.plonk import bitrev_indices
    from oracle import cbind as orc
    rng = np.random.default_rng(100 + log_n)
    c = rng.integers(0, P, (n_polys, 1 << log_n), dtype=np.uint64)
    got = _host(pv.lde(_dev(pv, c), rate_bits))
    want = orc.gl_lde(c, rate_bits)
    assert (got == want[:, bitrev_indices(log_n + rate_bits)]).all()
    # an explicit shift: the extension on another coset
    got2 = _host(pv.lde(_dev(pv, c[:1]), rate_bits, shift=7))
    assert (got2 == orc.gl_lde(c[:1], rate_bits, shift=7)[:, bitrev_indices(log_n + rate_bits)]).all()


def test_lde_large_is_evaluation_at_sampled_points(pv):
    """2^20 coefficients, rate 8 (the standard_recursion_config shape): sampled positions against Horner evaluation."""
    from blobstreamx_b200.plonk import bitrev_indices
    rng = np.random.default_rng(5)
    log_n, r = 20, 3
    c = rng.integers(0, P, (1, 1 << log_n), dtype=np.uint64)
    got = _host(pv.lde(_dev(pv, c), r))[0]
    g, wN = pv.coset_shift(), pv.root_of_unity(log_n + r)
    coeffs = [int(v) for v in c[0][::-1]]
    perm = bitrev_indices(log_n + r)
    for pos in (0, 1, 12345, (1 << 20) + 7, (1 << 23) - 1):
        x = g * pow(wN, int(perm[pos]), P) % P
        acc = 0
        for cj in coeffs:
            acc = (acc * x + cj) % P
        assert int(got[pos]) == acc, pos


@pytest.mark.parametrize("width,n_leaves,cap", [(3, 8, 1), (4, 16, 0), (11, 64, 2), (135, 256, 4), (135, 32, 5), (9, 1, 0)])
def test_merkle_caps(pv, width, n_leaves, cap):
    from oracle import cbind as orc
    rng = np.random.default_rng(width)
    data = rng.integers(0, 2**64, (width, n_leaves), dtype=np.uint64)
    d, capd = pv.merkle_caps(_dev(pv, data), cap)
    want = orc.gl_merkle(np.ascontiguousarray((data.astype(object) % P).astype(np.uint64).T), cap)
    assert (_host(d) == want).all()
    assert (_host(capd) == want[-(1 << cap):]).all()


@pytest.mark.parametrize("gate,p0,p1", [(0, 3, 0), (1, 2, 5), (2, 6, 0), (3, 32, 16), (4, 7, 0)])
def test_gate_quotient_matches_oracle(pv, gate, p0, p1):
    from oracle import cbind as orc
    rng = np.random.default_rng(gate)
    log_n, r = 6, 3
    rows = 1 << (log_n + r)
    nw, ncn = orc.gate_num_wires(gate, p0, p1), orc.gate_num_constraints(gate, p0, p1)
    w = rng.integers(0, P, (nw, rows), dtype=np.uint64)
    alphas = [int(v) for v in rng.integers(0, P, 2, dtype=np.uint64)]
    for na in (1, 2):
        ap, zh = pv.quotient_tables(alphas[:na], ncn, log_n, r)
        zhh = _host(zh)
        got = _host(pv.gate_quotient(gate, p0, p1, _dev(pv, w), ap, na, zh, log_n))
        want = orc.gl_quotient_combine(orc.gate_eval(gate, p0, p1, w), alphas[:na], np.repeat(zhh, 1 << log_n))
        assert (got == want).all()
    # the table: 1 / (x^n - 1) on block rev(q) of the coset
    g, wN = pv.coset_shift(), pv.root_of_unity(log_n + r)
    for q in range(1 << r):
        rq = int(f"{q:0{r}b}"[::-1], 2)
        xn = pow(g * pow(wN, q, P) % P, 1 << log_n, P)
        assert int(zhh[rq]) * ((xn - 1) % P) % P == 1


@pytest.mark.parametrize("arity_bits", [1, 3, 4])
def test_fri_fold(pv, arity_bits):
    from oracle import cbind as orc
    rng = np.random.default_rng(arity_bits)
    f = rng.integers(0, 2**64, (1 << 12, 2), dtype=np.uint64)
    beta = [int(v) for v in rng.integers(0, P, 2, dtype=np.uint64)]
    got = _host(pv.fri_fold(_dev(pv, f), arity_bits, beta))
    assert (got == orc.gl_fri_fold(f, arity_bits, beta)).all()


def test_trace_to_quotient_chain_on_device(pv):
    """A VALID U32Arithmetic trace (n = 2^10 rows, 114 wires) stays on the device: iNTT -> coset LDE (rate 8) -> Merkle cap
    of the extension + quotient.  Every constraint vanishes on the subgroup, so the alpha-combination is divisible by Z_H:
    the quotient (degree-4 constraints) must interpolate to a polynomial of degree < 3n -- its upper coefficients are zero.
    A corrupted trace must not pass that test."""
    import torch
    from blobstreamx_b200.plonk import bitrev_indices
    from oracle import cbind as orc
    rng = np.random.default_rng(21)
    log_n, r, gate, p0 = 10, 3, 0, 3
    n, N = 1 << log_n, 1 << (log_n + r)
    nw, ncn = orc.gate_num_wires(gate, p0, 0), orc.gate_num_constraints(gate, p0, 0)
    w = np.zeros((nw, n), np.uint64)
    for i in range(p0):
        w[6 * i:6 * i + 3] = rng.integers(0, 2**32, (3, n), dtype=np.uint64)
    w = orc.gate_witness(gate, p0, 0, w)
    assert not orc.gate_eval(gate, p0, 0, w).any()
    ap, zh = pv.quotient_tables([0x1234567, 0xABCDEF01], ncn, log_n, r)
    perm = bitrev_indices(log_n + r)
    g = pv.coset_shift()

    def chain(trace):
        coeffs = pv.ntt(_dev(pv, trace), inverse=True, natural_out=True)
        ext = pv.lde(coeffs, r)
        _, cap = pv.merkle_caps(ext, 4)
        q = pv.gate_quotient(gate, p0, 0, ext, ap, 2, zh, log_n)
        # back to coefficients: the buffer is bit-reversed, so un-permute, inverse transform, undo the coset shift
        qn = torch.empty_like(q)
        qn[:, torch.from_numpy(perm).to(pv.dev)] = q
        qc = _host(pv.ntt(qn, inverse=True, natural_out=True))
        return _host(ext), _host(cap), qc

    ext, cap, qc = chain(w)
    assert (ext == orc.gl_lde(orc.gl_ntt(w, inverse=True), r)[:, perm]).all()
    assert (cap == orc.gl_merkle(np.ascontiguousarray(ext.T), 4)[-16:]).all()
    assert not qc[:, 3 * n:].any() and qc[:, :3 * n].any()          # degree < 3n (coset scaling g^-j does not move zeros)
    bad = w.copy()
    bad[3, 17] ^= 1
    _, cap2, qc2 = chain(bad)
    assert qc2[:, 3 * n:].any() and (cap2 != cap).any()


@pytest.mark.parametrize("n_req", [1, 7, 400])
def test_sha256_trace_matches_oracle(pv, n_req):
    """SHA-256 execution trace on the device == the CPU restatement, whose columns tests/test_oracle_trace.py
    pins by recomputing every digest: fixed, variable and multi-chunk requests, padding rows zero."""
    import torch
    from oracle import cbind as orc
    from tests.test_oracle_trace import _requests
    rng = np.random.default_rng(n_req)
    bufs, offs, lens, kinds, _ = _requests(rng, n_req)
    hid = pv.ctx.hash_input_data(np.frombuffer(bufs, np.uint8), offs, lens, kinds)
    n = len(hid["padded_chunks"])
    log_rows = max(7, int(np.ceil(np.log2(64 * n))) + (1 if n_req == 7 else 0))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(pv.dev)
    got = pv.sha256_trace(dev(hid["padded_chunks"].view(np.int32)), dev(hid["end_bits"]), dev(hid["digest_bits"]), log_rows)
    want = orc.sha256_trace(hid["padded_chunks"], hid["end_bits"], hid["digest_bits"], log_rows)
    assert (_host(got) == want).all()
    # three circuits of the same schedule in one launch (bsx_sha256_trace_batch_dev), different words in each
    pcs = np.stack([hid["padded_chunks"], hid["padded_chunks"] ^ np.uint32(0x5A5A5A5A), rng.integers(0, 2**32, hid["padded_chunks"].shape, dtype=np.uint32)])
    ebs, dbs = np.stack([hid["end_bits"]] * 3), np.stack([hid["digest_bits"]] * 3)
    gotb = _host(pv.sha256_trace_batch(dev(pcs.view(np.int32)), dev(ebs), dev(dbs), log_rows))
    for c in range(3):
        assert (gotb[c] == orc.sha256_trace(pcs[c], hid["end_bits"], hid["digest_bits"], log_rows)).all()


def test_sha512_trace_matches_oracle(pv):
    """SHA-512 execution trace (the EdDSA accelerator of verify_skip: 100 requests of R ‖ A ‖ M) == the CPU restatement."""
    import torch
    from oracle import cbind as orc
    rng = np.random.default_rng(51)
    n_req = 100
    lens = (64 + rng.integers(100, 125, n_req)).astype(np.uint32)
    bufs = rng.integers(0, 256, 188 * n_req, dtype=np.uint8)
    offs = (np.arange(n_req + 1) * 188).astype(np.uint32)
    hid = pv.ctx.hash_input_data(bufs, offs, lens, np.ones(n_req, np.uint8), sha512=True)
    n = len(hid["padded_chunks"])
    assert n == 200
    log_rows = int(np.ceil(np.log2(80 * n)))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(pv.dev)
    got = pv.sha512_trace(dev(hid["padded_chunks"].view(np.int64)), dev(hid["end_bits"]), dev(hid["digest_bits"]), log_rows)
    assert (_host(got) == orc.sha512_trace(hid["padded_chunks"], hid["end_bits"], hid["digest_bits"], log_rows)).all()


def _ed_dev(pv, a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).to(pv.dev)


@pytest.mark.parametrize("log_rows,lanes", [(12, 0), (13, 1), (13, 8)])
def test_ed25519_trace_matches_oracle(pv, log_rows, lanes):
    """Ed25519 scalar-multiplication trace (EdDSA accelerator) on the device == oracle/ed_trace.py (pinned by
    tests/test_oracle_ed_trace.py): edge scalars (0, 1, 2^256 - 1, l), the identity and the order-2 point, random
    multiplications, s*G and h*A of a mocha-4 fixture signature, padding multiplications; k * P out of the same call."""
    import json, os
    from oracle import ed_trace as T
    from oracle import pyoracle as po
    from tests.test_oracle_ed_trace import _cases, _fixture_muls, _points_with_small_xy, pack
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mocha4.json")) as f:
        golden = json.load(f)
    pv.ctx.set_tunable("ED_TRACE_LANES", lanes)        # chain kernel: by batch size / one thread / eight lanes per multiplication
    rng = np.random.default_rng(8 + log_rows + lanes)
    scalars, points = _cases(rng, 2)
    if lanes == 8:      # curve points with x * y = 5, 8, 10, 16, 17: remainders below 19 in products (the quotient's corner case)
        corner = [p for t, p in _points_with_small_xy(19)]
        scalars, points = scalars + [int.from_bytes(rng.bytes(32), "little") | 1 for _ in corner], points + corner
    fs, fp, fw = _fixture_muls(golden, "10000", 1)
    scalars, points = scalars + fs, points + fp
    want, results = T.ed25519_trace(scalars, points, log_rows)
    assert results[-2:] == fw
    sc, pt = pack(scalars, points)
    n = len(scalars)
    got, res = pv.ed25519_trace(_ed_dev(pv, sc.reshape(n, 32)), _ed_dev(pv, pt.reshape(n, 64)), log_rows)
    assert (_host(got) == want).all()
    assert res.cpu().numpy().tobytes() == b"".join(po.ed_point_bytes(p) for p in results)
    # the two halves called apart (what a caller overlapping batches does) give the same table
    import torch
    d_sc, d_pt = _ed_dev(pv, sc.reshape(n, 32)), _ed_dev(pv, pt.reshape(n, 64))
    scratch, out2 = pv.ed25519_trace_scratch(n), torch.zeros_like(got)
    pv.ed25519_trace_points(d_sc, d_pt, scratch)
    pv.ed25519_trace_rows(d_sc, d_pt, scratch, log_rows, out2)
    assert torch.equal(out2, got)
    pv.ctx.set_tunable("ED_TRACE_LANES", 0)


def test_ed25519_trace_of_a_signature_batch_closes_on_the_witness_records(pv):
    """At batch size: 2 x 256 signatures' multiplications (s, G) and (h, A), taken from the 576-byte records of
    bsx_ed25519_batch (h at 64, s*G at 136, A at 200, h*A at 296; parity-green vs the oracle elsewhere) -> 2^17 rows.  Size-
    independent properties read off the trace: k * P equals the record's s*G / h*A, the last row's selected accumulator is k * P,
    rows chain (temp' = dbl, acc' = bit ? sum : acc) across the whole table, every limb is 16 bits, padding rows are flagged."""
    import torch
    from blobstreamx_b200 import synthetic as S
    from oracle import pyoracle as po
    n_sig = 256
    pks, sigs, msgs, lens, act = S.ed25519_batch_inputs(n_sig, inactive_every=50)      # 5 lanes run on the DUMMY signature
    rec = pv.ctx.ed25519_batch(pks, sigs, msgs, lens, act)
    assert (rec[:, 520] == 0xF).all() and (act == 0).sum() == 5
    G = np.frombuffer(po.ed_point_bytes(po.G), np.uint8)
    s_used = np.where(act[:, None] != 0, sigs[:, 32:64], np.frombuffer(po.DUMMY_SIGNATURE[32:], np.uint8)[None, :])
    scalars = np.empty((2 * n_sig, 32), np.uint8)
    points = np.empty((2 * n_sig, 64), np.uint8)
    scalars[0::2], points[0::2] = s_used, G
    scalars[1::2], points[1::2] = rec[:, 64:96], rec[:, 200:264]
    want = np.empty((2 * n_sig, 64), np.uint8)
    want[0::2], want[1::2] = rec[:, 136:200], rec[:, 296:360]
    # the operands gathered on the device from signatures + records (bsx_ed25519_trace_operands_dev) are those
    d_sc, d_pt = pv.ed25519_trace_operands(_ed_dev(pv, sigs), _ed_dev(pv, rec), _ed_dev(pv, act))
    assert (d_sc.cpu().numpy() == scalars).all() and (d_pt.cpu().numpy() == points).all()
    log_rows = 18                                                    # 131 072 real rows + as many padding rows: 3.2 GB
    tr, res = pv.ed25519_trace(d_sc, d_pt, log_rows)
    pv.ctx.set_tunable("ED_TRACE_LANES", 1)                          # the other chain kernel: identical table
    tr1, res1 = pv.ed25519_trace(_ed_dev(pv, scalars), _ed_dev(pv, points), log_rows)
    pv.ctx.set_tunable("ED_TRACE_LANES", 0)
    assert torch.equal(tr1, tr) and torch.equal(res1, res)
    del tr1
    assert (res.cpu().numpy() == want).all()
    n_real = 2 * n_sig * 256
    assert int(tr.max()) < 65536 and int(tr.min()) >= 0
    assert bool((tr[1, :n_real] == 1).all()) and bool((tr[1, n_real:] == 0).all())
    bit = tr[0] == 1
    last = tr[3] == 1
    inner = ~last
    inner[-1] = False
    nxt = torch.roll(tr, -1, dims=1)
    sum_xy = torch.cat([tr[68 + 92 * 6:68 + 92 * 6 + 16], tr[68 + 92 * 7:68 + 92 * 7 + 16]])
    dbl_xy = torch.cat([tr[68 + 92 * 14:68 + 92 * 14 + 16], tr[68 + 92 * 15:68 + 92 * 15 + 16]])
    sel = torch.where(bit[None, :], sum_xy, tr[36:68])
    assert bool((nxt[4:36][:, inner] == dbl_xy[:, inner]).all())
    assert bool((nxt[36:68][:, inner] == sel[:, inner]).all())
    fin = sel[:, last][:, :2 * n_sig].cpu().numpy().astype(np.uint16).T.copy()      # [mul, 32 limbs] -> 64 bytes little-endian
    assert (fin.view(np.uint8) == want).all()
    # the first 2^14 rows (64 multiplications) and the padding tail against the C restatement
    from oracle import cbind as orc
    ctr, _ = orc.ed25519_trace(scalars[:64], points[:64], 15, threads=orc.max_threads())
    assert (_host(tr[:, :1 << 14]) == ctr[:, :1 << 14]).all() and (_host(tr[:, -256:]) == ctr[:, -256:]).all()


def test_verify_skip_records_to_ed25519_trace(pv):
    """The EC accelerator of a verify_skip circuit end to end: the reference's 10000 -> 10500 fixture (2 real signatures, 98
    DUMMY lanes) through bsx_verify_skip, the 200 ScalarMul operands gathered on the device straight from the 240-byte
    validator records (signature at 32, lane flag at 236, strides BSX_VAL_IN_BYTES) and the 576-byte witness records, the
    2^16-row trace, and k * P of every multiplication equal to the records' s*G / h*A; the first real and the first DUMMY
    signature's rows against the oracle."""
    import json, os
    import torch
    from blobstreamx_b200 import inputs as I
    from oracle import ed_trace as T
    from oracle import pyoracle as po
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mocha4.json")) as f:
        golden = json.load(f)
    k = I.get_skip_inputs(golden["headers"]["10000"], golden["validators"]["10000"], golden["headers"]["10500"],
                          golden["commits"]["10500"], golden["validators"]["10500"])
    got = pv.ctx.verify_skip([k])
    assert got["fail"][0] == 0
    rec = np.ascontiguousarray(got["ed"][0])                                   # [100, 576]
    val = np.ascontiguousarray(np.asarray(k["target"]["validators"], np.uint8).reshape(100, 240))
    active = val[:, 236]
    assert 0 < int(active.sum()) < 100
    d_val, d_rec = _ed_dev(pv, val), _ed_dev(pv, rec)
    d_sc, d_pt = pv.ed25519_trace_operands(d_val[:, 32:96], d_rec, d_val[:, 236])
    sc, pt = d_sc.cpu().numpy(), d_pt.cpu().numpy()
    dummy_s = np.frombuffer(po.DUMMY_SIGNATURE[32:], np.uint8)
    assert (sc[0::2] == np.where(active[:, None] != 0, val[:, 64:96], dummy_s[None, :])).all() and (sc[1::2] == rec[:, 64:96]).all()
    assert (pt[0::2] == np.frombuffer(po.ed_point_bytes(po.G), np.uint8)).all() and (pt[1::2] == rec[:, 200:264]).all()
    tr, res = pv.ed25519_trace(d_sc, d_pt, 16)
    want = np.empty((200, 64), np.uint8)
    want[0::2], want[1::2] = rec[:, 136:200], rec[:, 296:360]
    assert (res.cpu().numpy() == want).all()
    assert bool((tr[1, :51200] == 1).all()) and bool((tr[1, 51200:] == 0).all())
    # the whole 2^16-row table against the C restatement (oracle/ed25519.c), then two signatures against the Python one
    from oracle import cbind as orc
    want_tr, want_res = orc.ed25519_trace(sc, pt, 16, threads=orc.max_threads())
    assert (want_res == want).all() and (_host(tr) == want_tr).all()
    first_real, first_dummy = int(np.argmax(active != 0)), int(np.argmax(active == 0))
    for i in (first_real, first_dummy):
        ks = [int.from_bytes(sc[2 * i + j].tobytes(), "little") for j in range(2)]
        ps = [(int.from_bytes(pt[2 * i + j, :32].tobytes(), "little"), int.from_bytes(pt[2 * i + j, 32:].tobytes(), "little")) for j in range(2)]
        w, _ = T.ed25519_trace(ks, ps, 9)
        assert (_host(tr[:, 512 * i:512 * (i + 1)]) == w).all()


def test_ed25519_trace_to_commitment_on_device(pv):
    """The Ed25519 trace stays on the device into the first commitment of a STARK over it: its 1540 columns are polynomials in
    the layout the transforms read -- iNTT -> coset LDE (rate 2) -> Poseidon Merkle cap over leaves of 1540 elements, every
    stage against the CPU restatements (the trace against oracle/ed25519.c).  Which blow-up and cap height starkyx's CurtaConfig
    uses is not on disk: rate 2 / cap height 4 here, parity unpinned like the rest of the prover loops."""
    from blobstreamx_b200.plonk import bitrev_indices
    from oracle import cbind as orc
    from oracle import pyoracle as po
    rng = np.random.default_rng(77)
    scalars = rng.integers(0, 256, (3, 32), dtype=np.uint8)
    pts = [po.G, po.ed_mul(12345, po.G), po.ed_mul(int.from_bytes(rng.bytes(32), "little"), po.G)]
    points = np.frombuffer(b"".join(po.ed_point_bytes(p) for p in pts), np.uint8).reshape(3, 64).copy()
    log_rows = 10
    tr, _ = pv.ed25519_trace(_ed_dev(pv, scalars), _ed_dev(pv, points), log_rows)
    want_tr, _ = orc.ed25519_trace(scalars, points, log_rows, threads=orc.max_threads())
    assert (_host(tr) == want_tr).all()
    coeffs = pv.ntt(tr, inverse=True, natural_out=True)
    ext = pv.lde(coeffs, 1)
    _, cap = pv.merkle_caps(ext, 4)
    want_ext = orc.gl_lde(orc.gl_ntt(want_tr, inverse=True), 1)[:, bitrev_indices(log_rows + 1)]
    assert (_host(ext) == want_ext).all()
    assert (_host(cap) == orc.gl_merkle(np.ascontiguousarray(want_ext.T), 4)[-16:]).all()

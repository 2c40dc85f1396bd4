"""Edge cases and error behaviour of the C ABI on the GPU: empty batches are no-ops, invalid arguments are refused with
a status and a message (the reference panics; the library never aborts), single-element batches work."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from blobstreamx_b200 import lib
    c = lib.Context(0)
    yield c
    c.close()


def test_empty_batches_are_noops(ctx):
    from oracle import cbind as orc
    l0 = ctx.launch_count
    assert ctx.sha256_batch(np.zeros(0, np.uint8), np.array([0], np.uint32)).shape == (0, 32)
    assert ctx.sha512_batch(np.zeros(0, np.uint8), np.array([0], np.uint32)).shape[0] == 0
    assert ctx.ed25519_batch(np.zeros((0, 32), np.uint8), np.zeros((0, 64), np.uint8), np.zeros((0, 124), np.uint8)).shape[0] == 0
    assert ctx.gl_poseidon_batch(np.zeros(0, np.uint64), np.array([0], np.uint32)).shape == (0, 4)
    assert ctx.header_trees(np.zeros((0, 512), np.uint8)).shape == (0, 32)
    nw = orc.gate_num_wires(orc.GATE_U32_ARITHMETIC, 3, 0)
    assert ctx.gl_gate_eval(orc.GATE_U32_ARITHMETIC, 3, 0, np.zeros((nw, 0), np.uint64)).size == 0
    from blobstreamx_b200 import inputs as I
    assert ctx.encode_headers(np.zeros(0, I.HEADER_FIELDS_DTYPE)).shape == (0, 512)
    assert ctx.validator_records(np.zeros(0, I.COMMIT_DTYPE), np.zeros((0, 100), I.COMMIT_SIG_DTYPE), 100)["validators"].shape == (0, 100, 240)
    assert ctx.present_on_trusted(np.zeros((0, 100), I.COMMIT_SIG_DTYPE), [], np.zeros((0, 100), I.COMMIT_SIG_DTYPE), [],
                                  np.zeros((0, 100, 240), np.uint8)).size == 0
    assert ctx.launch_count == l0          # nothing was launched
    # a zero-length message and a zero-length hash input are real work, not empty batches
    assert ctx.sha256_batch(np.zeros(0, np.uint8), np.array([0, 0], np.uint32))[0].tobytes().hex().startswith("e3b0c442")
    h = ctx.gl_poseidon_batch(np.zeros(0, np.uint64), np.array([0, 0], np.uint32))
    assert (h == orc.poseidon_batch(np.zeros(0, np.uint64), np.array([0, 0], np.uint32))).all()


def test_invalid_arguments_are_refused(ctx):
    from blobstreamx_b200 import lib
    from blobstreamx_b200 import synthetic as S
    m, _, _ = S.header_range_inputs(2, 4, None, with_skip=False)
    a = (m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers, m.end_headers, np.array([m.start_block], np.uint64),
         m.start_header, np.array([m.end_block], np.uint64), m.end_header)
    with pytest.raises(lib.BsxError, match="invalid argument"):
        ctx.prove_data_commitment(1, 3, 4, *a)                       # n_jobs must be a power of two (mapreduce tree)
    with pytest.raises(lib.BsxError, match="invalid argument"):
        ctx._call("bsx_sha256_batch", C.c_void_p(0), C.c_void_p(0), C.c_uint32(4), C.c_void_p(0))   # NULL buffers
    with pytest.raises(lib.BsxError, match="invalid argument"):
        w, c = np.zeros((4, 4), np.uint64), np.zeros((4, 4), np.uint64)
        ctx._call("bsx_gl_gate_eval", C.c_uint32(99), C.c_uint32(3), C.c_uint32(0), w.ctypes.data_as(C.c_void_p), C.c_uint32(4),
                  c.ctypes.data_as(C.c_void_p))                       # unknown gate
    with pytest.raises(lib.BsxError, match="invalid argument"):    # neither records nor hash fields requested
        ctx._call("bsx_validator_records", C.c_uint32(1), C.c_uint32(4), w.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                  C.c_void_p(0), C.c_void_p(0), C.c_void_p(0), C.c_void_p(0), c.ctypes.data_as(C.c_void_p))
    with pytest.raises(lib.BsxError, match="invalid argument"):    # hash fields come as a group of three arrays
        ctx._call("bsx_validator_records", C.c_uint32(1), C.c_uint32(4), w.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p),
                  C.c_void_p(0), c.ctypes.data_as(C.c_void_p), C.c_void_p(0), C.c_void_p(0), c.ctypes.data_as(C.c_void_p))
    with pytest.raises(lib.BsxError, match="invalid argument"):
        ctx._call("bsx_encode_headers", C.c_uint32(3), C.c_void_p(0), C.c_void_p(0))
    # the ctx stays usable after an error
    got = ctx.prove_data_commitment(1, 2, 4, *a)
    assert got["fail"][0] == 0


def test_single_element_batches(ctx):
    from oracle import cbind as orc
    from blobstreamx_b200 import synthetic as S
    pks, sigs, msgs, lens, active = S.ed25519_batch_inputs(1)
    assert (ctx.ed25519_batch(pks, sigs, msgs, lens, active) == orc.ed25519_batch(pks, sigs, msgs, lens, active)).all()
    m, _, _ = S.header_range_inputs(1, 1, None, with_skip=False)          # one job, one header
    got = ctx.prove_data_commitment(1, 1, 1, m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers, m.end_headers,
                                    np.array([m.start_block], np.uint64), m.start_header, np.array([m.end_block], np.uint64), m.end_header)
    want = orc.prove_data_commitment(1, 1, m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers, m.end_headers,
                                     m.start_block, m.start_header, m.end_block, m.end_header)
    assert got["fail"][0] == want["fail"] == 0
    assert got["data_commitments"][0].tobytes() == want["data_commitment"]
    assert (got["map_digests"][0] == want["map_digests"]).all()

"""Every build of the Ed25519 witness kernels, selected in-process through bsx_set_tunable, against the oracle on the
same stress inputs (valid, corrupted, garbage, edge encodings), and the co-run build of the pipelined host path
reached through the real entry point: one bsx_header_range call with more than 16 384 signatures."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

# (name, tunables): all force the one-thread-per-signature kernel except "quad"
BUILDS = [
    ("quad-lane three-stage path", {"ED_MODE": 1}),
    ("quad-lane three-stage path, integer limbs", {"ED_MODE": 1, "ED_FP64": 0}),
    ("compact, 216 registers", {"ED_MODE": 2, "ED_INLINE": 0, "ED_REGS": 0, "ED_FP64": 0}),
    ("inlined point arithmetic", {"ED_MODE": 2, "ED_INLINE": 1, "ED_REGS": 0, "ED_FP64": 0}),
    ("capped at 192 registers", {"ED_MODE": 2, "ED_REGS": 1, "ED_FP64": 0}),
    ("168 registers (6 CTAs/SM)", {"ED_MODE": 2, "ED_OCC": 6, "ED_FP64": 0}),
    ("128 registers (8 CTAs/SM, the co-run build)", {"ED_MODE": 2, "ED_OCC": 8, "ED_FP64": 0}),
    ("FP64-pipe field arithmetic", {"ED_MODE": 2, "ED_FP64": 1}),
    ("FP64-pipe field arithmetic, 128 registers", {"ED_MODE": 2, "ED_FP64": 1, "ED_OCC": 8}),
]


@pytest.fixture(scope="module")
def cases():
    from oracle import cbind as orc
    from tests._ed_cases import stress_inputs
    a = stress_inputs(1500, 2100, base_n=500)
    return a, orc.ed25519_batch(*a, threads=8)


# the per-key table path of the FP64 build (h*A from tabulated windows of each distinct public key, found on the device)
KEYED = [
    ("per-key tables", {"ED_MODE": 2, "ED_FP64": 1, "ED_KEYTAB": 1}),
    ("per-key tables, 128 registers", {"ED_MODE": 2, "ED_FP64": 1, "ED_KEYTAB": 1, "ED_OCC": 8}),
    ("per-key tables, 168 registers", {"ED_MODE": 2, "ED_FP64": 1, "ED_KEYTAB": 1, "ED_OCC": 6}),
    ("per-key tables, two signatures per thread", {"ED_MODE": 2, "ED_FP64": 1, "ED_KEYTAB": 1, "ED_PAIR": 1}),
    ("per-key tables, two signatures per thread, 128 registers", {"ED_MODE": 2, "ED_FP64": 1, "ED_KEYTAB": 1, "ED_PAIR": 1, "ED_OCC": 8}),
    ("per-key tables by the repeat rule", {"ED_MODE": 2, "ED_FP64": 1}),
]


@pytest.fixture(scope="module")
def few_keys():
    """~450 distinct keys (100 signers, one-bit corruptions of pk / R / s / message, garbage and edge encodings), each
    valid signer ~6 times: below the 1024-key capacity, so ED_KEYTAB = 1 really takes the table path; tiled 40 times
    (36 000 signatures, > 16 uses per key) the default rule takes it too."""
    from oracle import cbind as orc
    from tests._ed_cases import stress_inputs
    a = stress_inputs(600, 300, base_n=100)
    return a, orc.ed25519_batch(*a, threads=8)


@pytest.mark.parametrize("name,tun", KEYED, ids=[b[0] for b in KEYED])
def test_per_key_table_path_matches_oracle(few_keys, name, tun):
    from blobstreamx_b200 import lib
    a, want = few_keys
    reps = 40 if "ED_KEYTAB" not in tun else 1
    if "128 registers" in name and tun.get("ED_PAIR"):      # an odd count: the last thread of the pairing has one signature
        a, want = tuple(np.ascontiguousarray(x[:-1]) for x in a), want[:-1]
    if reps > 1:
        a = tuple(np.ascontiguousarray(np.concatenate([x] * reps)) for x in a)
        want = np.concatenate([want] * reps)
    with lib.Context(0) as ctx:
        for k, v in tun.items():
            ctx.set_tunable(k, v)
        got = ctx.ed25519_batch(*a)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (name, bad[:10])
    flags = want[:, 520]
    assert (flags == 0xF).sum() > 300 * reps and (flags != 0xF).sum() > 300 * reps


def test_per_key_tables_overflow_falls_back(cases):
    """More distinct keys than the table holds (the 3 600-key stress sample): the device-side verdict sends the whole
    batch down the general path, whatever ED_KEYTAB asks for."""
    from blobstreamx_b200 import lib
    a, want = cases
    with lib.Context(0) as ctx:
        for k, v in {"ED_MODE": 2, "ED_FP64": 1, "ED_KEYTAB": 1}.items():
            ctx.set_tunable(k, v)
        got = ctx.ed25519_batch(*a)
    assert (got == want).all()


@pytest.mark.parametrize("name,tun", BUILDS, ids=[b[0] for b in BUILDS])
def test_every_ed25519_build_matches_oracle(cases, name, tun):
    from blobstreamx_b200 import lib
    a, want = cases
    with lib.Context(0) as ctx:
        for k, v in tun.items():
            ctx.set_tunable(k, v)
            assert ctx.get_tunable(k) == v
        got = ctx.ed25519_batch(*a)
    bad = np.nonzero((got != want).any(axis=1))[0]
    assert bad.size == 0, (name, bad[:10])
    flags = want[:, 520]
    assert (flags == 0xF).sum() > 900 and (flags != 0xF).sum() > 1500      # the sample really has both kinds


def test_unknown_tunable_is_rejected():
    from blobstreamx_b200 import lib
    with lib.Context(0) as ctx:
        with pytest.raises(lib.BsxError):
            ctx.set_tunable("NO_SUCH_KNOB", 1)


def test_header_range_host_path_above_quad_threshold():
    """170 ranges x 100 validators = 17 000 signatures in ONE bsx_header_range call: the skip half takes the
    thread-per-signature kernel in its co-run (128-register) build beside the chunked map pipeline.  Every output of
    every range against the oracle (two distinct chains, tiled)."""
    import bench
    from blobstreamx_b200 import lib, synthetic as S
    from oracle import cbind as orc
    J, B, R = 2, 4, 170
    vs = S.ValidatorSet.make(S.SEED)
    sets = [S.header_range_inputs(J, B, None, start=2_000_000 + 1000 * r, seed=S.SEED + 7 + r, valset=vs) for r in range(2)]
    ms, skips = [x[0] for x in sets], [x[1] for x in sets]
    mm = bench.tile_ranges(ms, R)
    mm["n_jobs"], mm["batch"] = J, B
    with lib.Context(0) as ctx:
        got = ctx.header_range([skips[r % 2] for r in range(R)], mm)
    assert not got["fail"].any() and not got["skip"]["fail"].any()
    for d in range(2):
        m = ms[d]
        w = orc.prove_data_commitment(J, B, m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers, m.end_headers,
                                      m.start_block, m.start_header, m.end_block, m.end_header)
        ws = orc.verify_skip(skips[d], threads=8)
        for r in range(d, R, 2):
            assert (got["map_digests"][r] == w["map_digests"]).all(), r
            assert (got["reduce_nodes"][r] == w["reduce_nodes"]).all() and got["data_commitments"][r].tobytes() == w["data_commitment"]
            assert (got["skip"]["sha256_digests"][r] == ws["sha256_digests"]).all(), r
            assert (got["skip"]["ed"][r] == ws["ed"]).all(), r

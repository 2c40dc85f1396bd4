"""Synthetic workloads (blobstreamx_b200.synthetic) are valid witnesses for the oracle's circuits. CPU only."""
import numpy as np
import pytest

from blobstreamx_b200 import synthetic as S
from oracle import cbind as orc


@pytest.fixture(scope="module")
def valset():
    return S.ValidatorSet.make()


@pytest.mark.parametrize("J,B,nblk", [(4, 8, None), (4, 8, 19), (2, 32, 33), (8, 4, 1)])
def test_header_range_synthetic(valset, J, B, nblk):
    m, skip, chain = S.header_range_inputs(J, B, nblk, valset=valset, with_skip=False)
    r = orc.prove_data_commitment(J, B, m.dh_leaf, m.dh_aunts, m.lb_leaf, m.lb_aunts, m.start_headers, m.end_headers,
                                  m.start_block, m.start_header, m.end_block, m.end_header, threads=2)
    assert r["fail"] == 0
    n = m.end_block - m.start_block
    from oracle import pyoracle as po
    dhs = [bytes.fromhex(chain.headers[i]["data_hash"]) for i in range(n)]
    assert r["data_commitment"] == po.data_commitment(dhs, m.start_block)


def test_skip_synthetic(valset):
    m, skip, chain = S.header_range_inputs(2, 4, valset=valset)
    r = orc.verify_skip(skip, threads=orc.max_threads())
    assert r["fail"] == 0
    assert (r["ed"][:, 520] == 0xF).all()
    # nil + absent votes and a tampered signature
    commit = S.make_commit(chain, chain.start + 8, absent=(3,), nil=(90,))  # a nil vote inside the 1/3 prefix trips the reference quirk at conversion.rs:222-231
    from blobstreamx_b200 import inputs as I
    k = I.get_skip_inputs(chain.headers[0], valset.validators, chain.headers[-1], commit, valset.validators)
    assert k["target"]["validators"][3, 236] == 0 and k["target"]["validators"][90, 236] == 0
    assert orc.verify_skip(k, threads=orc.max_threads())["fail"] == 0
    k["target"]["validators"][0, 40] ^= 1
    assert orc.verify_skip(k, threads=orc.max_threads())["fail"] & 1

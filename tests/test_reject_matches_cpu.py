"""registration::RejectMatchesRanSaPC as the library implements it -- host code by design (one sequential chain of
std::default_random_engine draws), so it runs without a GPU -- against the oracle, which is pinned to the compiled reference."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracleapi


def _case(n_src=800, n_tgt=900, n_pairs=700, seed=2):
    rng = np.random.default_rng(seed)
    src = rng.uniform(-1, 1, (n_src, 3)).astype(np.float32)
    R = np.linalg.qr(rng.normal(size=(3, 3)))[0].astype(np.float32)
    tgt = np.concatenate([src @ R.T + 0.1, rng.uniform(-1, 1, (n_tgt - n_src, 3)).astype(np.float32)]).astype(np.float32)
    pairs = np.stack([rng.integers(0, n_src, n_pairs), rng.integers(0, n_tgt, n_pairs)], 1).astype(np.int32)
    good = rng.random(n_pairs) < 0.4
    pairs[good, 1] = pairs[good, 0]            # 40 % true correspondences
    return src, tgt, pairs


@pytest.mark.parametrize("rounds,cand,diff", [(1, 4, 0.1), (3, 4, 0.1), (3, 2, 0.05), (2, 8, 0.01)])
def test_library_rejection_equals_the_oracle(rounds, cand, diff):
    from onepiece_b200 import registration as reg
    src, tgt, pairs = _case()
    engine = reg.DefaultRandomEngine()
    kept = pairs
    for _ in range(rounds):
        kept = reg.RejectMatchesRanSaPC(src, tgt, engine, kept, cand, diff)
    want = oracleapi.reject_matches(src, tgt, pairs, rounds, cand, diff)
    assert np.array_equal(kept, want) and 0 < len(kept) < len(pairs)
    assert engine.state.value != 1


def test_rejection_argument_errors():
    from onepiece_b200 import capi, registration as reg
    src, tgt, pairs = _case()
    assert len(reg.RejectMatchesRanSaPC(src, tgt, reg.DefaultRandomEngine(), pairs[:0])) == 0
    bad = pairs.copy()
    bad[3, 1] = len(tgt)
    with pytest.raises(capi.OpbError):
        reg.RejectMatchesRanSaPC(src, tgt, reg.DefaultRandomEngine(), bad)
    with pytest.raises(capi.OpbError):
        reg.RejectMatchesRanSaPC(src, tgt, reg.DefaultRandomEngine(0), pairs)

"""The engine's reformulated math (bit matrices, index compares, per-chain
cut-offs; tests/engine_model.py) against the literal oracle, on CPU."""
import numpy as np
import pytest

import engine_model as em
import golden_specs as gs
import oracle as orc
from util import assert_same

NAMES = ["g1_m4_n2000_s1_k50", "g1_m4_n2000_s3_k7", "g1_m7_n3000_s4_k11_stake",
         "g2_m16_n12000_s1_k500", "g3_m16_n6000_s1_k700", "g1_m33_n6000_s7_k640"]


@pytest.mark.parametrize("name", NAMES)
def test_model_matches_oracle(name):
    tr, K, stake = gs.make_trace(name)
    o = orc.run_oracle(tr, K, stake)
    m = em.run_model(tr, K, stake)
    assert_same(o, m, what=name)
    assert np.array_equal(o["oracle"].can_see(), m["can_see"])

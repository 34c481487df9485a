"""The CPU research prototypes under tools/research: the integer model of the reference's compensated sum
must track the floating-point loop exactly, and its chunked evaluation must reproduce it.  Small sizes: seconds."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "research"))


def test_integer_model_tracks_the_floating_point_loop():
    import kahan_automaton as ka
    assert ka.check(n_trials=36, seed=11) == 0


def test_chunked_evaluation_is_exact():
    import kahan_automaton as ka
    import kahan_chunks as kc
    rng = np.random.default_rng(5)
    for kind in range(3):
        n = 700
        ts = [rng.integers(0, 256, n), rng.choice([255, 254, 0, 1, 200], n), rng.choice([255, 128, 127, 64], n)][kind]
        s = c = 0.0
        k = 0
        while s < 4.0:
            s, c = ka.kahan_fp(ts[k:k + 1], s, c)
            k += 1
        A0 = ka.to_A(s, c)
        s_fp, c_fp = ka.kahan_fp(ts[k:], s, c)
        got, _ = kc.chained(ts[k:], A0, 128)
        assert got == ka.to_A(s_fp, c_fp)
        assert ka.rn53(got)[0] * 2.0 ** -ka.G == s_fp


def test_anchored_four_branch_segments_are_exact():
    import kahan_segments as ks
    assert ks.check(seed=9, trials=8) == 0

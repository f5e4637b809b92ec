"""CPU: host-side policies of the engine that need no device -- the width of the prediction tiles and the switch of
the predictive-variance term from triangular solves to the explicit inverse of the factor."""
from projected_lmc_b200 import engine as E


def test_tile_points_is_a_multiple_of_128_between_128_and_8192():
    tp = E.LatentEngine.tile_points
    assert tp(4, 44544, budget_bytes=6 << 30) == 4480            # 6 GiB / (4 * 44544 * 8) = 4519 -> 35 * 128
    assert tp(16, 20096, budget_bytes=6 << 30) == 2432
    assert tp(16, 20096, budget_bytes=40 << 30) == 8192          # capped at the width the GEMMs like
    assert tp(64, 131072, budget_bytes=1 << 20) == 128           # never below one tile


def test_prediction_inverts_the_factor_once_the_points_reach_half_n(monkeypatch):
    calls = []
    monkeypatch.setattr(E.ops, "trtri", lambda L, dinv, cfg=None: calls.append("trtri"))
    eng = E.LatentEngine()
    eng._cfg = (None, None)
    st = {"n": 1000, "L": None, "dinv": None}
    assert eng.predict_inverse == "auto"
    assert eng._maybe_invert(st, 200) is False and eng._maybe_invert(st, 200) is False and calls == []
    assert eng._maybe_invert(st, 100) is True and calls == ["trtri"] and st["inverted"]      # 500 >= 1000 / 2
    assert eng._maybe_invert(st, 1) is True and calls == ["trtri"]                           # inverted once
    eng.predict_inverse = False
    st2 = {"n": 10, "L": None, "dinv": None}
    assert eng._maybe_invert(st2, 10 ** 6) is False and calls == ["trtri"]
    eng.predict_inverse = True
    assert eng._maybe_invert(st2, 1) is True and calls == ["trtri", "trtri"]

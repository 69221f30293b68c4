"""CPU: Elo bookkeeping restated from games/algos/elo.py:45-71 (key order, swap, accumulation) and the rating fit."""
import math

import pytest

from self_play_reinforcement_learning_b200 import elo


def test_merge_results_key_convention_and_accumulation():
    shelf = {}
    assert elo.merge_results(shelf, "modelb", "modela", {"wins": 6, "draws": 1, "losses": 3}) == "modelb__modela"
    assert shelf["modelb__modela"] == {"wins": 6, "draws": 1, "losses": 3}
    # the same pairing reported from the other side lands in the same key, mirrored (elo.py:59-66)
    assert elo.merge_results(shelf, "modela", "modelb", {"wins": 2, "draws": 0, "losses": 8}) == "modelb__modela"
    assert shelf["modelb__modela"] == {"wins": 14, "draws": 1, "losses": 5}
    with pytest.raises(AssertionError):
        elo.merge_results(shelf, "a_b", "c", {"wins": 0, "draws": 0, "losses": 0})     # elo.py:47-48
    with pytest.raises(AssertionError):
        elo.merge_results(shelf, "x", "x", {"wins": 0, "draws": 0, "losses": 0})


def test_fit_elo_recovers_known_differences():
    # 75 % score against the anchor <=> +400*log10(3) Elo
    shelf = {"strong__random": {"wins": 750, "draws": 0, "losses": 250}}
    r = elo.fit_elo(shelf, ["random", "strong"])
    assert r["random"] == 0 and abs(r["strong"] - 400 * math.log10(3)) < 1e-6
    # draws are half points; three models, consistent ratings 0 / 200 / 300
    def score(d):
        return 1 / (1 + 10 ** (-d / 400))
    n = 100000
    shelf = {"random__mida": {"wins": 0, "draws": int(2 * n * score(-200)), "losses": n - int(2 * n * score(-200))},
             "top__mida": {"wins": int(n * score(100)), "draws": 0, "losses": n - int(n * score(100))},
             "top__random": {"wins": int(n * score(300)), "draws": 0, "losses": n - int(n * score(300))}}
    r = elo.fit_elo(shelf, ["random", "mida", "top"], anchor_model="random", anchor_elo=0)
    assert abs(r["mida"] - 200) < 0.5 and abs(r["top"] - 300) < 0.5
    # anchor offset moves everything
    r2 = elo.fit_elo(shelf, ["random", "mida", "top"], anchor_model="random", anchor_elo=1000)
    assert abs(r2["top"] - r["top"] - 1000) < 1e-6


def test_model_database_registry():
    db = elo.ModelDatabase("tictactoe")
    db.add_model("random", "random")
    with pytest.raises(ValueError):
        db.add_model("random", "random")
    assert db.env == 1 and db.get_model("random") == "random" and db.elos() == {}

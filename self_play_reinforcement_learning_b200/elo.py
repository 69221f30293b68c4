"""Elo glue: the reference's ``Elo`` bookkeeping (games/algos/elo.py:14-161) and ``ModelDatabase``
registry (games/algos/model_database.py:11-77) over the batched GPU engine.

The reference plays each pairing by building a whole ``SelfPlayScheduler`` (worker processes, two InferenceWorkers) and
calling ``compare_models`` (elo.py:73-91); here one ``BatchedSelfPlay`` plays all ``num_games`` head-to-head games of a
pairing at once with both networks native on the GPU (``nets.TwoTowerEvaluator``) or against the device-side hard-coded
players.  Kept: result-key convention ``"{larger}__{smaller}"`` with wins/draws/losses from the larger name's view
(elo.py:45-71), accumulation over repeated comparisons, ``compare_models(*names)`` over all pairs, ``calculate_elo`` with an
anchor model fixed at ``anchor_elo`` and expected score ``q1 / (q1 + q2)``, ``q = 10 ** (rating / 400)``, draws worth 0.5
(elo.py:93-161).  Changed: ``calculate_elo`` minimises the same binary cross-entropy with deterministic full-batch Newton
steps instead of 100 000 noisy SGD steps (elo.py:110-129); shelves are plain dicts (persistence is out of scope).
"""
import itertools
import math

import numpy as np

from .scheduler import parse_results
from .selfplay import BatchedSelfPlay

HARDCODED = ("random", "lookahead")     # ModelContainer(policy_gen=Random / OneStepLookahead), hardcoded_players.py:8-56


class ModelDatabase:
    """model_database.py:11-77 without the shelve files: name -> nn.Module, or one of HARDCODED."""

    def __init__(self, game="connect4", env=None):
        self.game, self.env = game, env if env is not None else {"connect4": 0, "tictactoe": 1}[game]
        self.model_shelf, self.result_shelf, self.elo_value_shelf = {}, {}, {}

    def add_model(self, name, model):
        if name in self.model_shelf:
            raise ValueError("Model name already in use")      # model_database.py:66-68
        self.model_shelf[name] = model

    def get_model(self, name):
        return self.model_shelf[name]

    def elos(self):
        return dict(self.elo_value_shelf)


def merge_results(result_shelf, model_1, model_2, new_results):
    """Elo._compare bookkeeping (elo.py:45-71): ``new_results`` are from model_1's view."""
    assert model_1 != model_2 and "_" not in model_1 and "_" not in model_2
    if model_1 > model_2:
        key, ordered = f"{model_1}__{model_2}", dict(new_results)
    else:
        key = f"{model_2}__{model_1}"
        ordered = {"wins": new_results["losses"], "draws": new_results["draws"], "losses": new_results["wins"]}
    old = result_shelf.get(key, {"wins": 0, "draws": 0, "losses": 0})
    result_shelf[key] = {k: ordered[k] + old[k] for k in ("wins", "draws", "losses")}
    return key


def fit_elo(result_shelf, models, anchor_model="random", anchor_elo=0.0, iterations=200, prior_games=0.0):
    """Maximum-likelihood ratings for E[score of a vs b] = 1 / (1 + 10 ** ((r_b - r_a) / 400)) with the anchor fixed; a draw
    counts as half a win and half a loss (result_map elo.py:149).  ``prior_games`` virtual draws per listed pairing keep
    ratings finite for perfect records."""
    free = [m for m in models if m != anchor_model]
    index = {m: i for i, m in enumerate(free)}
    k = math.log(10.0) / 400.0
    pairs = []
    for key, res in result_shelf.items():
        a, b = key.split("__")
        if a not in models or b not in models:
            continue
        n = res["wins"] + res["draws"] + res["losses"] + prior_games
        if n:
            pairs.append((a, b, res["wins"] + 0.5 * res["draws"] + 0.5 * prior_games, n))
    r = np.zeros(len(free))
    for _ in range(iterations):
        g, h = np.zeros(len(free)), np.zeros((len(free), len(free)))
        for a, b, s, n in pairs:
            ra = r[index[a]] if a in index else anchor_elo
            rb = r[index[b]] if b in index else anchor_elo
            e = 1.0 / (1.0 + math.exp(-k * (ra - rb)))
            d, w = k * (s - n * e), k * k * n * e * (1.0 - e)
            for m, sign in ((a, 1.0), (b, -1.0)):
                if m in index:
                    g[index[m]] += sign * d
                    for m2, sign2 in ((a, 1.0), (b, -1.0)):
                        if m2 in index:
                            h[index[m], index[m2]] += sign * sign2 * w
        if not len(free) or np.abs(g).max() < 1e-12:
            break
        step = np.linalg.solve(h + 1e-12 * np.eye(len(free)), g)
        r = r + np.clip(step, -400.0, 400.0)
    out = {m: float(r[index[m]]) for m in free}
    out[anchor_model] = float(anchor_elo)
    return out


class Elo:
    ELO_CONSTANT = 400

    def __init__(self, model_database, iterations=400, n_games_per_gpu=1024, seed=0, net="tower"):
        self.model_database, self.iterations, self.n_games_per_gpu, self.seed, self.net = model_database, iterations, n_games_per_gpu, seed, net
        self._comparisons = 0

    def compare_all(self):
        self.compare_models(*list(self.model_database.model_shelf.keys()))

    def compare_models(self, *args, num_games=100):
        for model_1, model_2 in itertools.combinations(args, 2):
            self._compare(model_1, model_2, num_games=num_games)

    def _compare(self, model_1, model_2, num_games=100):
        return merge_results(self.model_database.result_shelf, model_1, model_2, self._get_results(model_1, model_2, num_games))

    def _get_results(self, model_1, model_2, num_games=100):
        """elo.py:73-91: num_games evaluation games, model_1 = the policy, model_2 = the evaluation policy, alternating who
        starts; returns wins/draws/losses from model_1's view."""
        shelf = self.model_database.model_shelf
        a, b = shelf[model_1], shelf[model_2]
        flip = False
        if isinstance(a, str) and a in HARDCODED:
            if isinstance(b, str):
                raise ValueError("at least one side of a comparison must be a network")
            a, b, flip = b, a, True                      # the searching side must be the network
        self._comparisons += 1
        target = num_games + (num_games & 1)
        G = min(self.n_games_per_gpu, target)
        kw = dict(env=self.model_database.env, n_games=G, sims=self.iterations, evaluate=True, update=False, games_target=target,
                  seed=self.seed + 104729 * self._comparisons, eval_cache=True)   # same games, fewer network passes
        if isinstance(b, str):
            sp = BatchedSelfPlay(a, net=self.net, opponent=b, **kw)
        else:
            sp = BatchedSelfPlay(a, net=self.net, evaluation_network=b, **kw)
        _, results = sp.play_games()
        sp.close()
        _, breakdown = parse_results(results[:num_games])
        res = {s: breakdown["first"][s] + breakdown["second"][s] for s in ("wins", "draws", "losses")}
        if flip:
            res = {"wins": res["losses"], "draws": res["draws"], "losses": res["wins"]}
        return res

    def calculate_elo(self, anchor_model="random", anchor_elo=0):
        models = list(self.model_database.model_shelf.keys())
        elos = fit_elo(self.model_database.result_shelf, models, anchor_model, anchor_elo, prior_games=1.0)
        self.model_database.elo_value_shelf["elo"] = elos
        return elos

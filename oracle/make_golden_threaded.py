"""Golden vectors for the THREADED search (MCTreeSearch(thread_count=K) behind an InferenceProxy, mcts.py:328-331): the
unmodified reference run under the cooperative round-robin schedule of oracle/ref_harness._coop_search (one legal interleaving
of its K threads) -> tests/golden/threaded.json.  TEST INFRASTRUCTURE ONLY.

    python -m oracle.make_golden_threaded
"""
import json
import os

import numpy as np

from . import ref_harness as rh
from . import spec

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    rng = np.random.default_rng(4242)
    searches = []
    cases = [(0, 60, 2, ()), (0, 200, 4, ()), (0, 401, 4, ()), (1, 100, 4, ()), (1, 64, 3, ()), (0, 3, 4, ()), (0, 150, 8, ()),
             (0, 300, 4, [(3, 1), (2, -1), (3, 1), (3, -1)]), (1, 200, 4, [(4, 1), (0, -1), (8, 1), (2, -1)]),
             (0, 250, 4, [(0, 1), (1, -1), (0, 1), (1, -1), (0, 1)])]
    for i, (game, sims, K, prefix) in enumerate(cases):
        A = spec.GAME_DIMS[game][2]
        noise = rng.dirichlet([1.0] * A)
        kw = dict(seed=300 + i, game_uid=11 * i, net_seed=i, strong_play=(i == 9))
        r = rh.run_search(game, sims, noise=noise, prefix=prefix, threads=K, **kw)
        searches.append(dict(name=f"thr{i}", game=game, sims=sims, threads=K, prefix=[list(p) for p in prefix], noise=[float(x).hex() for x in noise],
                             n=r["n"].tolist(), w=[float(x).hex() for x in r["w"]], valid=r["valid"].tolist(), root_n=r["root_n"],
                             root_w=float(r["root_w"]).hex(), q=float(r["q"]).hex(), player=r["player"], **kw))
        print("search", i, game, sims, K, r["n"].tolist(), r["root_n"])
    eps = []
    specs = [(0, 60, 4, False, False), (0, 90, 4, True, False), (0, 50, 2, False, True), (1, 80, 4, False, False), (1, 40, 4, True, True),
             (0, 120, 8, True, False)]
    for i, (game, sims, K, swap, ev) in enumerate(specs):
        A = spec.GAME_DIMS[game][2]
        table = rng.dirichlet([1.0 if i % 2 else 0.3] * A, size=(2, 22))
        opp_seed = 55 if ev else None
        uid = 3000 + 2 * i + int(swap)
        r = rh.run_episode(game, sims, seed=70 + i, game_uid=uid, swap_sides=swap, evaluate=ev, noise_table=table, net_seed=i,
                           net_seed_opp=opp_seed, threads=K)
        eps.append(dict(game=game, sims=sims, threads=K, swap=swap, evaluate=ev, seed=70 + i, game_uid=uid, net_seed=i, net_seed_opp=opp_seed,
                        noise_table=[[[float(x).hex() for x in row] for row in t] for t in table], reward=r["reward"],
                        moves=[dict(tree=m["tree"], ply=m["ply"], action=m["action"], n=m["n"], w=[float(x).hex() for x in m["w"]],
                                    root_n=m["root_n"], root_w=float(m["root_w"]).hex()) for m in r["moves"]],
                        records=[dict(state=rec["state"].tolist(), actual_val=rec["actual_val"],
                                      tree_probs=[float(x).hex() for x in rec["tree_probs"]], q=float(rec["q"]).hex()) for rec in r["records"]],
                        final_state=r["final_state"].tolist()))
        print("episode", i, game, sims, K, swap, ev, "reward", r["reward"], "plies", len(r["moves"]))
    with open(os.path.join(OUT, "threaded.json"), "w") as f:
        json.dump(dict(searches=searches, episodes=eps), f)
    print("written", os.path.join(OUT, "threaded.json"))


if __name__ == "__main__":
    main()

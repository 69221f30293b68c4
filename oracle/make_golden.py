"""Writes tests/golden/*.npz by running the UNMODIFIED reference (/root/reference) under the
oracle/ref_harness.py hooks.  Run here (the only place the reference exists):

    python -m oracle.make_golden

The fixtures are the pins for the C oracle (tests/test_oracle_golden.py) and, through it, for the
CUDA engine.  Inputs are fully determined by the parameters stored next to each result.
"""
import json
import os

import numpy as np

from . import ref_harness as rh
from . import spec

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def random_actions(game, n, T, rng, legal_only):
    """Action lists: legal random playouts (Connect4) padded with extra moves after the end
    (exercises GameOver / full-column ValueError), or unconstrained (TicTacToe no-op quirk)."""
    W, H, A = spec.GAME_DIMS[game]
    acts = rng.integers(0, A, size=(n, T)).astype(np.int32)
    if legal_only:
        for g in range(n):
            heights = np.zeros(W, int)
            for t in range(T):
                if rng.random() < 0.97:  # mostly legal, sometimes deliberately a full column
                    legal = np.flatnonzero(heights < H)
                    if len(legal):
                        acts[g, t] = rng.choice(legal)
                if heights[acts[g, t]] < H:
                    heights[acts[g, t]] += 1
    return acts


def adversarial_c4():
    rows = [
        [0, 1, 0, 1, 0, 1, 0, 2, 2],                    # vertical win on col 0, then GameOver steps
        [0, 0, 1, 1, 2, 2, 3, 4, 4],                    # horizontal win bottom row
        [0, 1, 1, 2, 2, 3, 2, 3, 3, 6, 3],              # rising diagonal
        [6, 5, 5, 4, 4, 3, 4, 3, 3, 0, 3],              # falling diagonal
        [3, 3, 3, 3, 3, 3, 3, 3],                       # full column then ValueError twice
        [0, 6, 0, 6, 0, 6, 1, 5, 1, 5, 1, 5, 0],        # no win yet
    ]
    # a full-board draw: columns filled in an order known to produce no 4-in-a-row
    draw = []
    order = [0, 1, 2, 3, 4, 5, 6]
    pattern = {0: "XOXOXO", 1: "XOXOXO", 2: "OXOXOX", 3: "OXOXOX", 4: "XOXOXO", 5: "XOXOXO", 6: "OXOXOX"}
    # build a legal alternating move order realising `pattern` (X = +1 moves first)
    heights = [0] * 7
    turn = "X"
    for _ in range(42):
        for c in order:
            if heights[c] < 6 and pattern[c][heights[c]] == turn:
                draw.append(c)
                heights[c] += 1
                turn = "O" if turn == "X" else "X"
                break
        else:
            break
    rows.append(draw + [0])
    T = 44
    acts = -np.ones((len(rows), T), np.int32)
    for i, r in enumerate(rows):
        acts[i, :len(r)] = r
    return acts


def main():
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)

    # ---- env fixtures
    a_c4 = np.concatenate([adversarial_c4(), random_actions(spec.GAME_CONNECT4, 300, 44, rng, True)])
    fp = np.where(np.arange(len(a_c4)) % 2 == 0, 1, -1).astype(np.int8)
    ref = rh.ref_env_playout(spec.GAME_CONNECT4, a_c4, fp)
    np.savez_compressed(os.path.join(OUT, "env_connect4.npz"), actions=a_c4, first_player=fp, **ref)
    a_t = random_actions(spec.GAME_TICTACTOE, 400, 12, rng, False)
    fp = np.where(np.arange(len(a_t)) % 2 == 0, 1, -1).astype(np.int8)
    ref = rh.ref_env_playout(spec.GAME_TICTACTOE, a_t, fp)
    np.savez_compressed(os.path.join(OUT, "env_tictactoe.npz"), actions=a_t, first_player=fp, **ref)

    # ---- search fixtures (fresh root / after a prefix of play_action calls)
    cases = []
    A7 = [1.0 / 7] * 7

    class Uniform:
        def __init__(self, A): self.A = A
        def to(self, *a, **k): return self
        def __call__(self, s, player=1): return [1.0 / self.A] * self.A, 0.0

    def add_search(name, game, sims, **kw):
        net_kind = kw.pop("net_kind", "hash")
        network = Uniform(spec.GAME_DIMS[game][2]) if net_kind == "uniform" else None
        r = rh.run_search(game, sims, network=network, **kw)
        kw2 = {k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in kw.items()}
        cases.append(dict(name=name, game=game, sims=sims, net_kind=net_kind, params=kw2,
                          n=r["n"].tolist(), w=[float(x).hex() for x in r["w"]], valid=r["valid"].tolist(),
                          root_n=r["root_n"], root_w=float(r["root_w"]).hex(), q=float(r["q"]).hex(), player=r["player"]))

    # SURVEY.md Appendix B known-answer vectors #1-#5
    add_search("B1", 0, 50, tie_mode=0, net_kind="uniform")
    add_search("B2", 0, 800, tie_mode=0, net_kind="uniform")
    add_search("B3", 1, 100, tie_mode=0, net_kind="uniform")
    add_search("B4", 0, 200, tie_mode=0, net_kind="uniform", prefix=[(3, 1), (3, -1)] * 3)
    add_search("B5", 0, 200, tie_mode=0, net_kind="uniform", prefix=[(0, 1), (0, -1), (1, 1), (1, -1), (2, 1), (2, -1)])
    for i, (game, sims) in enumerate([(0, 50), (0, 200), (0, 800), (0, 800), (1, 100), (1, 400), (0, 1600)]):
        A = spec.GAME_DIMS[game][2]
        noise = rng.dirichlet([1.0] * A)
        add_search(f"hash{i}", game, sims, seed=100 + i, game_uid=7 * i, net_seed=i, noise=noise)
    add_search("hash_prefix", 0, 300, seed=5, game_uid=1, net_seed=2, noise=rng.dirichlet([0.15] * 7),
               prefix=[(3, 1), (2, -1), (3, 1), (3, -1)])
    add_search("hash_strong", 0, 400, seed=6, game_uid=2, net_seed=3, strong_play=True,
               prefix=[(0, 1), (1, -1), (0, 1), (1, -1), (0, 1)])
    add_search("ttt_late", 1, 200, seed=8, game_uid=3, net_seed=4, prefix=[(4, 1), (0, -1), (8, 1), (2, -1)])
    with open(os.path.join(OUT, "search.json"), "w") as f:
        json.dump(cases, f)

    # ---- episode fixtures
    eps = []
    specs = [(0, 50, False, False, False), (0, 80, True, False, False), (0, 120, False, True, False),
             (0, 60, True, True, False), (0, 100, False, False, True), (1, 100, False, False, False),
             (1, 60, True, False, False), (1, 50, False, True, False), (1, 40, True, True, False),
             (0, 200, False, False, False), (0, 30, True, False, False), (0, 30, False, False, False)]
    for i, (game, sims, swap, ev, strong) in enumerate(specs):
        A = spec.GAME_DIMS[game][2]
        alpha = 1.0 if i % 3 else 0.15
        table = rng.dirichlet([alpha] * A, size=(2, 22))
        opp_seed = 77 if ev else None
        uid = 1000 + 2 * i + int(swap)  # engine convention: swap_sides == game index odd (self_play_parallel.py:252)
        r = rh.run_episode(game, sims, seed=40 + i, game_uid=uid, swap_sides=swap, evaluate=ev,
                           noise_table=table, net_seed=i, net_seed_opp=opp_seed, strong_play=strong)
        eps.append(dict(game=game, sims=sims, swap=swap, evaluate=ev, strong_play=strong, seed=40 + i,
                        game_uid=uid, net_seed=i, net_seed_opp=opp_seed,
                        noise_table=[[[float(x).hex() for x in row] for row in t] for t in table],
                        reward=r["reward"],
                        moves=[dict(tree=m["tree"], ply=m["ply"], action=m["action"], n=m["n"],
                                    w=[float(x).hex() for x in m["w"]], root_n=m["root_n"], root_w=float(m["root_w"]).hex())
                               for m in r["moves"]],
                        records=[dict(state=rec["state"].tolist(), actual_val=rec["actual_val"],
                                      tree_probs=[float(x).hex() for x in rec["tree_probs"]], q=float(rec["q"]).hex())
                                 for rec in r["records"]],
                        final_state=r["final_state"].tolist()))
        print("episode", i, game, sims, swap, ev, "reward", r["reward"], "plies", len(r["moves"]))
    with open(os.path.join(OUT, "episodes.json"), "w") as f:
        json.dump(eps, f)

    # ---- evaluation games against the reference's hard-coded players (general/hardcoded_players.py)
    vs = []
    k = 0
    for game in (0, 1):
        for kind in (spec.OPP_LOOKAHEAD, spec.OPP_RANDOM):
            for swap in (False, True):
                for rep in range(2):
                    uid = 5000 + 2 * k + int(swap)
                    k += 1
                    r = rh.run_episode_vs_hardcoded(game, 40, kind, seed=900 + k, game_uid=uid, swap_sides=swap, net_seed=k)
                    vs.append(dict(game=game, kind=kind, swap=swap, sims=40, seed=900 + k, game_uid=uid, net_seed=k, reward=r["reward"],
                                   moves=[[m["tree"], m["ply"], m["action"]] for m in r["moves"]], final_state=r["final_state"].tolist()))
    with open(os.path.join(OUT, "episodes_vs_hardcoded.json"), "w") as f:
        json.dump(vs, f)
    # ---- network fixtures: the reference's own classes, random init under a fixed seed, fp32 CPU
    import torch
    rh._import_reference()
    from games.general.modules import ResidualTower
    from games.tictactoe.modules import ConvNetTicTacToe
    nets = {}
    torch.set_num_threads(1)
    for name, ctor, seed, (W, Hh) in [("tower20", lambda: ResidualTower(7, 6, 7, num_blocks=20), 0, (7, 6)),
                                      ("tower2", lambda: ResidualTower(7, 6, 7, num_blocks=2), 3, (7, 6)),
                                      ("tower_ttt", lambda: ResidualTower(3, 3, 9, num_blocks=3), 4, (3, 3)),
                                      ("convttt", lambda: ConvNetTicTacToe(3, 3, 9), 1, (3, 3))]:
        torch.manual_seed(seed)
        net = ctor().eval()
        x = torch.from_numpy(rng.integers(-1, 2, size=(24, W, Hh)).astype(np.int64))
        x[0] = 0
        with torch.no_grad():
            p, v = net.forward(x)
        nets[name + "_x"] = x.numpy().astype(np.int8)
        nets[name + "_policy"] = p.numpy()
        nets[name + "_value"] = v.numpy()
        nets[name + "_wsum"] = np.array([sum(float(t.double().abs().sum()) for t in net.state_dict().values())])
        nets[name + "_seed"] = np.array([seed])
    # ---- loss fixture: MCTreeSearch.loss (mcts.py:234-252) on a fixed batch with the reference's own code
    import games.algos.mcts as ref_mcts
    torch.manual_seed(3)
    net = ResidualTower(7, 6, 7, num_blocks=2).eval()
    lrng = np.random.default_rng(77)
    batch = []
    for _ in range(16):
        st = torch.from_numpy(lrng.integers(-1, 2, size=(7, 6)).astype(np.int64))
        pr = torch.from_numpy(lrng.dirichlet([1.0] * 7).astype(np.float32))
        batch.append(ref_mcts.Move(st, torch.tensor(float(lrng.integers(-1, 2))), pr, torch.tensor(float(lrng.uniform(-1, 1)))))
    holder = type("H", (), {})()
    holder.network, holder.q_average = net, True
    with torch.no_grad():
        lval = float(ref_mcts.MCTreeSearch.loss(holder, batch))
    nets["loss_states"] = np.stack([b.state.numpy() for b in batch]).astype(np.int8)
    nets["loss_probs"] = np.stack([b.tree_probs.numpy() for b in batch])
    nets["loss_val"] = np.array([float(b.actual_val) for b in batch], np.float32)
    nets["loss_q"] = np.array([float(b.q) for b in batch], np.float32)
    nets["loss_value"] = np.array([lval])
    np.savez_compressed(os.path.join(OUT, "nets.npz"), **nets)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()

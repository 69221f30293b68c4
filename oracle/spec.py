"""Shared, exactly reproducible test-input specification (TEST INFRASTRUCTURE ONLY).

Nothing in this file restates the reference; it defines the *injected inputs* that the
reference (via ``oracle/ref_harness.py`` hooks), the C oracle (``oracle/spx_oracle.c``) and
the CUDA engine (``csrc/spx_rng.cuh``) all consume so that "identical network outputs,
Dirichlet noise and tie-breaking" (BASELINE.json north_star) is a checkable statement:

* ``splitmix64`` counter-based stream keyed by (seed, game_uid, tree, purpose, ply, sim,
  depth, idx) -> uniform double in [0,1) with 53 random bits.  Integer-only until the final
  exact ``* 2**-53`` so CPU and GPU agree bit for bit.
* the synthetic "hash net": a pure function (own bitboard, opp bitboard) -> (policy f32[A],
  value f32) built from one correctly-rounded fp32 division per output.
* bitboard encodings of the two games' boards.

The reference call sites these stand in for: ``np.random.rand(A)`` mcts.py:355,
``np.random.dirichlet`` mcts.py:50, ``np.random.choice`` mcts.py:280, and the network
callable ``network(s, player)`` mcts.py:168,316.
"""
import numpy as np

M64 = (1 << 64) - 1

PURPOSE_TIE = 0      # tie-break noise, one value per (sim, depth, action)
PURPOSE_GAMMA = 1    # Dirichlet/Gamma variate stream (device-generated noise mode)
PURPOSE_ACTION = 2   # the single uniform that np.random.choice consumes per move
PURPOSE_OPPONENT = 3 # random.choice of the hard-coded opponents (hardcoded_players.py:29,49): index = int(u * n)

OPP_MCTS, OPP_LOOKAHEAD, OPP_RANDOM = 0, 1, 2

GAME_CONNECT4 = 0
GAME_TICTACTOE = 1

GAME_DIMS = {GAME_CONNECT4: (7, 6, 7), GAME_TICTACTOE: (3, 3, 9)}  # (W, H, A)


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M64
    return z ^ (z >> 31)


def rng_u64(seed, game_uid, tree, purpose, ply, sim, depth, idx):
    h = splitmix64(seed & M64)
    h = splitmix64(h ^ (game_uid & M64))
    h = splitmix64(h ^ ((tree & 0xFF) | ((purpose & 0xFF) << 8) | ((ply & 0xFFFF) << 16)))
    h = splitmix64(h ^ ((sim & 0xFFFFFFFF) | ((depth & 0xFFFFFFFF) << 32)))
    h = splitmix64(h ^ (idx & M64))
    return h


def rng_uniform(seed, game_uid, tree, purpose, ply, sim, depth, idx):
    """Uniform double in [0,1): top 53 bits * 2**-53 (exact)."""
    return (rng_u64(seed, game_uid, tree, purpose, ply, sim, depth, idx) >> 11) * (1.0 / 9007199254740992.0)


# ----------------------------------------------------------------------------- boards
def board_to_bits(board, game):
    """int board [W,H] in {-1,0,1} -> (own, opp) python ints.

    Connect4: bit = col*7 + row (7 bits per column, bit 6 of each column is a never-set
    sentinel).  TicTacToe: bit = x*3 + y (== the action index, tictactoe_env.py:39-40).
    """
    W, H, _ = GAME_DIMS[game]
    stride = 7 if game == GAME_CONNECT4 else 3
    own = opp = 0
    b = np.asarray(board)
    for c in range(W):
        for r in range(H):
            v = int(b[c, r])
            if v == 1:
                own |= 1 << (c * stride + r)
            elif v == -1:
                opp |= 1 << (c * stride + r)
    return own, opp


def bits_to_board(own, opp, game):
    W, H, _ = GAME_DIMS[game]
    stride = 7 if game == GAME_CONNECT4 else 3
    b = np.zeros((W, H), dtype=np.int64)
    for c in range(W):
        for r in range(H):
            k = c * stride + r
            if (own >> k) & 1:
                b[c, r] = 1
            elif (opp >> k) & 1:
                b[c, r] = -1
    return b


# ----------------------------------------------------------------------------- hash net
def hashnet(own, opp, n_actions, net_seed=0):
    """Synthetic deterministic 'network': (own, opp) bitboards in the NET frame ->
    (policy float32[A] summing to ~1, value float32 in [-0.4, 0.4))."""
    k = splitmix64((own & M64) ^ splitmix64(((opp & M64) + (net_seed & M64)) & M64))
    r = np.array([(splitmix64(k ^ (i + 1)) >> 48) + 1 for i in range(n_actions)], dtype=np.int64)
    tot = np.float32(int(r.sum()))
    p = r.astype(np.float32) / tot
    vv = np.float32(int(splitmix64(k ^ 0xFF) >> 48)) - np.float32(32768.0)
    v = np.float32(vv / np.float32(81920.0))
    return p.astype(np.float32), v

"""B200-native batched self-play engine: a drop-in for the AlphaZero MCTS hot path of
reubenvanammers/self_play_reinforcement_learning (games/algos/mcts.py driven by
games/algos/selfplayworker.py / self_play_parallel.py).

The compute path is libspx.so (hand-written sm_100a CUDA behind the C ABI in include/spx.h);
this package is the thin Python/PyTorch host that mirrors the reference's plugin interfaces.
"""
from ._lib import GAME_CONNECT4, GAME_TICTACTOE, GAME_DIMS, SpxError  # noqa: F401

__all__ = ["GAME_CONNECT4", "GAME_TICTACTOE", "GAME_DIMS", "SpxError"]

"""Stand-in for gym (only ``gym.spaces.Discrete(n).n`` is used: connect4env.py:16, mcts.py:151)."""
from . import spaces  # noqa: F401

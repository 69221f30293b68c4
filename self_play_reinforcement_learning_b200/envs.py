"""Batched environments: the games/general/base_env.py interface (step / valid_moves / reset /
set_state / get_state / max_moves / num_actions / variant_string) over n boards at once, computed by
the libspx bitboard kernels (spx_env_step, spx_env_valid_moves).

Semantics follow games/connect4/connect4env.py:29-101 and games/tictactoe/tictactoe_env.py:23-101:
Connect4 boards are indexed [col,row] with row 0 at the bottom; stepping a finished game raises
GameOver, a full column raises ValueError; a TicTacToe move on an occupied cell is a silent no-op.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import GAME_CONNECT4, GAME_TICTACTOE, check, lib


class GameOver(Exception):
    """Same name/meaning as games/general/base_env.py:4-5."""


def _bit_index(game, device):
    W, H, _ = _lib.GAME_DIMS[game]
    stride = 7 if game == GAME_CONNECT4 else 3
    idx = torch.arange(W, device=device)[:, None] * stride + torch.arange(H, device=device)[None, :]
    return idx  # [W,H] int64


def boards_to_bits(boards, game):
    """int boards [n,W,H] in {-1,0,1} -> int64 [n,2] (own, opp) bitboards (oracle/spec.py layout)."""
    idx = _bit_index(game, boards.device)
    one = torch.ones((), dtype=torch.int64, device=boards.device)
    weights = one << idx
    own = ((boards == 1).to(torch.int64) * weights).sum(dim=(1, 2))
    opp = ((boards == -1).to(torch.int64) * weights).sum(dim=(1, 2))
    return torch.stack([own, opp], dim=1).contiguous()


def bits_to_boards(bits, game):
    """int64 [n,2] bitboards -> int64 boards [n,W,H]."""
    idx = _bit_index(game, bits.device)
    own = (bits[:, 0, None, None] >> idx) & 1
    opp = (bits[:, 1, None, None] >> idx) & 1
    return own - opp


class BatchedEnv:
    game = None
    name = None

    def __init__(self, n=1, device=None, strict=True):
        if not torch.cuda.is_available():
            raise _lib.SpxError("batched envs run on the GPU only (no CPU fallback)")
        self.n = int(n)
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.width, self.height, self._A = _lib.GAME_DIMS[self.game]
        self.strict = strict
        self.reset()

    # -- BaseEnv interface ------------------------------------------------------------------
    def __call__(self):  # env_gen() -> fresh copy (base_env.py:9-10)
        other = type(self)(self.n, self.device, self.strict)
        other.bits.copy_(self.bits)
        other.episode_over.copy_(self.episode_over)
        return other

    def num_actions(self):
        return self._A

    def max_moves(self):
        return self.width * self.height

    def variant_string(self):
        return self.name

    def reset(self):
        self.bits = torch.zeros(self.n, 2, dtype=torch.int64, device=self.device)
        self.episode_over = torch.zeros(self.n, dtype=torch.uint8, device=self.device)
        self._reward = torch.zeros(self.n, dtype=torch.int8, device=self.device)
        self._valid = torch.zeros(self.n, dtype=torch.int16, device=self.device)
        self.last_status = torch.zeros(self.n, dtype=torch.int8, device=self.device)
        return self.board

    @property
    def board(self):
        return bits_to_boards(self.bits, self.game)

    def set_state(self, state):
        """state: int boards [n,W,H]; like the reference, episode_over is left untouched."""
        state = torch.as_tensor(state, device=self.device).reshape(self.n, self.width, self.height)
        self.bits = boards_to_bits(state, self.game)

    def get_state(self):
        return self.board, None

    def valid_moves(self):
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib().spx_env_valid_moves(self.game, self.n, self.bits.data_ptr(), self._valid.data_ptr(), stream), "spx_env_valid_moves")
        return self._unpack_valid()

    def _unpack_valid(self):
        v = self._valid.to(torch.int32) & 0xFFFF
        return ((v[:, None] >> torch.arange(self._A, device=self.device)[None, :]) & 1).bool()

    def step(self, action, player=1):
        """action: int or int tensor [n] (<0 skips that board); player: +1/-1 or int tensor [n].
        Returns (boards, reward int8[n], done bool[n], None)."""
        a = torch.as_tensor(action, device=self.device).to(torch.int32).expand(self.n).contiguous()
        p = torch.as_tensor(player, device=self.device).to(torch.int8).expand(self.n).contiguous()
        stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        check(lib().spx_env_step(self.game, self.n, self.bits.data_ptr(), self.episode_over.data_ptr(), a.data_ptr(), p.data_ptr(),
                                 self._reward.data_ptr(), self._valid.data_ptr(), self.last_status.data_ptr(), stream), "spx_env_step")
        if self.strict:
            st = self.last_status
            if bool((st == -1).any()):
                raise GameOver
            if bool((st == -2).any()):
                raise ValueError("move into a full column")
        return self.board, self._reward.clone(), self.episode_over.bool(), None


class Connect4Env(BatchedEnv):
    game = GAME_CONNECT4
    name = "connect4"


class TicTacToeEnv(BatchedEnv):
    game = GAME_TICTACTOE
    name = "tictactoe"


def game_id_of(env):
    """Maps a reference env class/instance (or one of ours) to a built-in game id via variant_string()
    (connect4env.py:97-101, tictactoe_env.py:93-101); non-default board sizes are rejected."""
    if isinstance(env, int):
        return env
    obj = env
    if isinstance(env, type):
        if issubclass(env, BatchedEnv):
            return env.game
        obj = env()
    name = obj.variant_string()
    if name == "connect4":
        return GAME_CONNECT4
    if name == "tictactoe":
        return GAME_TICTACTOE
    raise ValueError(f"unsupported env variant {name!r}: only default connect4 (7x6) and tictactoe (3x3x3) are built in")

"""MCTreeSearch: the reference's per-game ``Policy`` object (games/algos/mcts.py:116-394, contract in
games/general/base_model.py:10-76) backed by one libspx engine slot.

It is what ``ModelContainer(MCTreeSearch, policy_kwargs=dict(iterations=800, ...)).setup(network=..., env=...,
memory_queue=...)`` returns, so the reference's own ``SelfPlayer.play_episode`` (selfplayworker.py:164-224), ``elo.py`` and
``external_play.py`` loops can drive a GPU search against ANY other ``BasePlayer`` (a human, OneStepLookahead, another
network): the opposing moves are delivered to the engine with ``spx_set_external_actions``.  For throughput use
``selfplay.BatchedSelfPlay`` (thousands of games per GPU); this facade is the drop-in for the one-game-at-a-time API.

Kept API: ``reset(player)``, ``__call__(s) -> action``, ``play_action(a, player)``, ``push_to_queue(done=, r=)``,
``evaluate(bool)``, ``train(bool)``, ``load_state_dict``, ``state_dict``, ``loss``, ``update_from_memory``, attributes
``env`` (assignable), ``network``, ``iterations``, ``strong_play``, ``moves_played``, ``memory``.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib, nets
from ._lib import check, lib
from .engine import HashNetEvaluator, SelfPlayEngine
from .envs import game_id_of
from .scheduler import Memory, mcts_loss
from .selfplay import Move, records_to_moves

Move = Move  # re-export under the reference's name (mcts.py:17)
OPP_EXTERNAL = 3


class MCTreeSearch:
    def __init__(self, network, env, optim=None, memory_queue=None, iterations=100, temperature_cutoff=5, batch_size=64,
                 memory_size=200000, min_memory=20000, update_nn=True, starting_state_dict=None, thread_count=4,
                 strong_play=False, q_average=True, alpha=1, net="auto", seed=0, net_dtype=torch.bfloat16, noise_mode=2):
        self.iterations, self.network, self.env_gen, self.optim = iterations, network, env, optim
        self.game = game_id_of(env)
        self.W, self.H, self.actions = _lib.GAME_DIMS[self.game]
        self.env = env() if callable(env) else env
        self.alpha, self.strong_play, self.q_average = alpha, strong_play, q_average
        self.temperature_cutoff, self.batch_size, self.min_memory, self.update_nn = temperature_cutoff, batch_size, min_memory, update_nn
        # the reference searches with thread_count threads only behind an InferenceProxy (mcts.py:154,328); called directly -- as this
        # object is -- it is the sequential search.  BatchedSelfPlay / SelfPlayScheduler(search_threads=K) run the threaded search.
        self.thread_count = 1
        self.memory_queue, self.memory = memory_queue, Memory(memory_size)
        self.evaluating = False
        self.seed, self._net_kind, self._net_dtype, self._noise_mode = seed, net, net_dtype, noise_mode
        self._engine, self._engine_key, self._episodes = None, None, 0
        self.moves_played = 0
        if starting_state_dict:
            self.load_state_dict(starting_state_dict)
        self.reset()

    # ------------------------------------------------------------------ engine plumbing
    def _evaluator(self):
        kind = self._net_kind
        if kind == "auto":
            ok = isinstance(self.network, nets.ResidualTower) or hasattr(self.network, "residual_blocks")
            ok = ok and getattr(self.network.conv1, "out_channels", 0) == 128   # filter_factor 32; 7x6 and 3x3 boards are native
            kind = "tower" if ok else ("hash" if self.network is None else "torch")
            if isinstance(self.network, nets.ConvNetTicTacToe) and self.game == _lib.GAME_TICTACTOE and self.network.action_size == 9:
                kind = "tttnet"
        if kind == "tower":
            return nets.TowerEvaluator(self.network, self.game)
        if kind == "hash":
            return HashNetEvaluator(self.game, self.seed)
        if kind == "tttnet":
            return nets.TTTNetEvaluator(self.network)
        return nets.TorchNetEvaluator(self.network, self.game, dtype=self._net_dtype)

    def _ensure_engine(self):
        key = (bool(self.evaluating), bool(self.strong_play), float(self.alpha), int(self.iterations))
        if self._engine is None or key != self._engine_key:
            if self._engine is not None:
                self._engine.close()
            self._engine = SelfPlayEngine(self.game, 1, self.iterations, self._evaluator(), evaluate=self.evaluating,
                                          strong_play=self.strong_play, alpha=self.alpha, seed=self.seed, opponent_kind=OPP_EXTERNAL,
                                          slot_stride=2, max_sims_per_tick=16, move_log=True, noise_mode=self._noise_mode)
            self._engine_key = key
            self._status = torch.zeros(1, 6, dtype=torch.int32, device=self._engine.device)
            self._ext = torch.full((1,), -1, dtype=torch.int32, device=self._engine.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def _slot(self):
        check(lib().spx_slot_status(self._engine._h, self._status.data_ptr(), self._stream()), "spx_slot_status")
        st = self._status.cpu().numpy()[0]
        return dict(state=int(st[0]), ply=int(st[1]), own_moves=int(st[2]), own_action=int(st[3]), games_finished=int(st[4]), swap=int(st[5]))

    # ------------------------------------------------------------------ Policy API
    def reset(self, player=1):
        """mcts.py:166-174.  player == -1: the opponent moves first in this policy's frame (selfplayworker.py:175-176)."""
        self._ensure_engine()
        swap = 1 if player == -1 else 0
        self.game_index = 2 * self._episodes + swap
        self._episodes += 1
        check(lib().spx_restart(self._engine._h, self.game_index, self.game_index + 1, self._stream()), "spx_restart")
        self._engine._first = True
        self.moves_played = 0
        self.temp_memory = []
        return np.zeros((self.W, self.H), dtype=np.int64)

    def __call__(self, s=None):
        """_search_and_play (mcts.py:177-186): run ticks until this policy has produced its next move."""
        want = self.moves_played + 1
        for _ in range(self.iterations * 4 + 64):
            self._engine.run_ticks(8)
            st = self._slot()
            if st["own_moves"] >= want:
                break
            if st["state"] == 2:
                raise _lib.SpxError("MCTreeSearch.__call__: the game is already over")
        else:
            raise _lib.SpxError("MCTreeSearch.__call__: the engine did not produce a move (is it the opponent's turn?)")
        self.moves_played = want
        self._last_action = st["own_action"]
        return int(self._last_action)

    def play_action(self, action, player):
        """mcts.py:188-209.  player == +1: this policy's own move (already applied on the device); otherwise the opposing
        move, delivered to the engine (selfplayworker.py:221-224 calls both policies for every move)."""
        if player == 1:
            return
        self._ext.fill_(int(action))
        check(lib().spx_set_external_actions(self._engine._h, self._ext.data_ptr(), self._stream()), "spx_set_external_actions")
        self._engine.run_ticks(2)   # consume it (re-root, env step); a pending evaluation continues on the next call

    def push_to_queue(self, s=None, a=None, r=None, done=None, next_s=None):
        """mcts.py:225-232: at game end stamp every pending record with the result and hand it to the memory queue."""
        if not done:
            return
        for _ in range(64):
            if self._slot()["state"] == 2:     # the engine has seen the end of the game and flushed the records
                break
            self._engine.run_ticks(2)
        recs = self._engine.drain_records()
        recs = recs[recs["game_index"] == self.game_index]
        for m in records_to_moves(recs, self.game):
            m = m._replace(actual_val=torch.tensor(r).float())
            if self.memory_queue is not None:
                self.memory_queue.put(m)
            else:
                self.memory.add(m)
        self.temp_memory = []

    def pull_from_queue(self):
        while self.memory_queue is not None and not self.memory_queue.empty():
            self.memory.add(self.memory_queue.get())

    def update(self, s, a, r, done, next_s):
        self.push_to_queue(s, a, r, done, next_s)
        self.pull_from_queue()
        if self.ready:
            self.update_from_memory()

    @property
    def ready(self):
        return len(self.memory) >= self.min_memory and self.update_nn

    def loss(self, batch):
        return mcts_loss(self.network, batch, self.q_average)

    def update_from_memory(self):
        if len(self.memory) < self.batch_size:
            return
        loss = self.loss(self.memory.sample(self.batch_size))
        self.optim.zero_grad()
        loss.backward()
        self.optim.step()
        self._engine_key = None   # weights changed: rebuild the evaluator on the next reset

    def evaluate(self, evaluate_state=False):
        self.evaluating = evaluate_state

    def train(self, train_state=True):
        return self.network.train(train_state) if self.network is not None else None

    def load_state_dict(self, state_dict, target=False):
        self.network.load_state_dict(state_dict)
        self._engine_key = None

    def state_dict(self):
        return self.network.state_dict()

    def update_target_net(self):
        pass

    def deduplicate(self):
        """mcts.py:385-386: self.memory.deduplicate("state", ["actual_val", "tree_probs"], Move) when a memory is attached."""
        memory = getattr(self, "memory", None)
        if memory is not None and hasattr(memory, "deduplicate"):
            memory.deduplicate("state", ["actual_val", "tree_probs"], Move)

    def root_stats(self):
        return self._engine.root_stats(0)

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None

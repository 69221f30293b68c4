"""SelfPlayScheduler: the reference's orchestration layer (games/algos/self_play_parallel.py:44-379) over the batched
GPU engine -- same epoch structure, same result bookkeeping, same loss, with the worker processes, queues and
checkpoint-file weight hand-off replaced by one BatchedSelfPlay per GPU and an NCCL weight broadcast.

Row (f1)/(f2) of SURVEY.md 8: records flow engine -> device replay memory (replay.DeviceReplay, csrc/spx_replay.cu) ->
sampled training batch -> the hand-written SGD step (train.DeviceTrainer, csrc/spx_train.cu) without touching the host, on
rank 0 (the reference trains in one UpdateWorker, updateworker.py:141-149).  Networks the native step is not built for
(filter_factor != 32, other boards) and the host replay memory train with PyTorch autograd (trainer="torch").
"""
from collections import deque

import numpy as np
import torch

import datetime
import os

from . import checkpoint, nets, parallel
from .replay import DeviceReplay, loss_from_batch
from .envs import game_id_of
from .selfplay import BatchedSelfPlay, Move, records_to_moves, results_to_dicts


class Memory:
    """rl_utils/memory.py:8-30: bounded FIFO of Move tuples with uniform sampling without replacement."""

    def __init__(self, max_size=None):
        self.max_size = max_size
        self._buffer = deque(maxlen=max_size)
        self._dedup, self._dedup_pending = None, []

    def __setstate__(self, state):
        """Unpickling (checkpoint.load_memory): files of the reference carry max_size, _buffer and `deduplicator`
        (memory.py:9-12) -- a Deduplicator's counter / temp_queue are this class's running sums / pending list."""
        state = dict(state)
        d = state.pop("deduplicator", None)
        self._dedup, self._dedup_pending = None, []
        self.__dict__.update(state)
        if d is not None and getattr(d, "counter", None) is not None:
            self._dedup, self._dedup_pending = dict(d.counter), list(getattr(d, "temp_queue", ()))

    def __len__(self):
        return len(self._buffer)

    def add(self, experience):
        self._buffer.append(experience)
        if self._dedup is not None:
            self._dedup_pending.append(experience)

    def change_size(self, max_size):
        self.max_size = max_size
        self._buffer = deque(self._buffer, maxlen=max_size)

    def sample(self, batch_size):
        index = np.random.choice(np.arange(len(self._buffer)), size=batch_size, replace=False)
        return [self._buffer[i] for i in index]

    def reset(self):
        self._buffer = deque(maxlen=self.max_size)

    def deduplicate(self, key="state", values=("actual_val", "tree_probs"), named_tuple=Move, maxlen=None):
        """memory.py:47-94 for host-side Move tuples: one entry per distinct `key` ever added, `values` averaged (running sums in
        insertion order / count), first-seen order, the last `maxlen`.  Fields of `named_tuple` that are neither key nor value
        (Move.q at the reference's only call site, where create_memory therefore raises TypeError) are averaged too."""
        fields = getattr(named_tuple, "_fields", (key, *values))
        values = [f for f in fields if f != key]
        if self._dedup is None:
            self._dedup, self._dedup_pending = {}, list(self._buffer)
        for e in self._dedup_pending:
            k = getattr(e, key).detach().cpu().numpy().tobytes()
            c = self._dedup.get(k)
            if c is None:
                self._dedup[k] = {"count": 1, key: getattr(e, key), **{v: getattr(e, v) for v in values}}
            else:
                c["count"] += 1
                for v in values:
                    c[v] = c[v] + getattr(e, v)
        self._dedup_pending = []
        self._buffer = deque((named_tuple(**{key: c[key], **{v: c[v] / c["count"] for v in values}}) for c in self._dedup.values()), maxlen=maxlen)


class DeduplicatorState:
    """What a pickled rl_utils.memory.Deduplicator becomes when a reference memory file is loaded here (its attributes,
    no behaviour): Memory.__setstate__ takes over `counter` and `temp_queue`."""


def mcts_loss(network, batch, q_average=True):
    """MCTreeSearch.loss (mcts.py:234-252): MSE(mean) of the value against actual_val (+ q when q_average) plus the
    policy cross-entropy -sum(log p * tree_probs) / B."""
    s, actual_val, tree_probs, q = Move(*zip(*batch))
    dev = next(network.parameters()).device
    net_probs, predict_val = network.forward(torch.stack(s).to(dev))
    predict_val = predict_val.view(-1)
    target = torch.stack(actual_val).to(dev)
    if q_average:
        target = target + torch.stack(q).to(dev)
    value_loss = torch.nn.functional.mse_loss(predict_val, target.float())
    prob_loss = -(net_probs.log() * torch.stack(tree_probs).to(dev)).sum() / net_probs.size(0)
    return value_loss + prob_loss


def parse_results(reward_list):
    """self_play_parallel.py:302-327: (total_rewards, {"first": {...}, "second": {...}}); 'first' = swap_sides False."""
    breakdown = {}
    for j, start in enumerate(("first", "second")):
        sel = [r for r in reward_list if r["swap_sides"] == bool(j)]
        breakdown[start] = dict(wins=sum(r["reward"] == 1 for r in sel), draws=sum(r["reward"] == 0 for r in sel),
                                losses=sum(r["reward"] == -1 for r in sel))
    total_rewards = int(np.sum([r["reward"] for r in reward_list])) if reward_list else 0
    return total_rewards, breakdown


class SelfPlayScheduler:
    def __init__(self, network, env, evaluation_network=None, iterations=800, epoch_length=1500, initial_games=64,
                 evaluation_games=150, games_per_gpu=1024, memory_size=200000, batch_size=128, lr=0.01, momentum=0.9,
                 weight_decay=1e-4, updates_per_epoch=100, alpha=1.0, seed=0, net="tower", evaluation_opponent=None,
                 replay="device", max_memory_size=None, memory_step=0, deduplicate=False, save_dir=None, save_memory=True,
                 lr_patience=15, amp=None, trainer="auto", search_threads=1, eval_cache=True):
        """replay: "device" (records stay in HBM: DeviceReplay) or "host" (reference-style Memory of Move tuples).
        max_memory_size / memory_step: UpdateWorker's staggered buffer growth (updateworker.py:107-109).
        save_dir: as in the reference (self_play_parallel.py:56,263-267): every epoch rank 0 writes
        ``<save_dir>/<start_time>/model-<iso time>:<games>`` ({"model": state_dict}, updateworker.py:111-117) and, with
        ``save_memory``, the pickled Memory of Move tuples ``memory-<iso time>:<size>`` (previous file removed, :119-139);
        ``train_model(resume_model=, resume_memory=)`` picks up the newest files of the previous run (base_worker.py:26-62).
        (Pickling a full 200 000-record Memory of Move tuples takes about a minute, as it does in the reference;
        ``save_memory=False`` keeps only the model checkpoints.)
        search_threads: thread_count of the policies' MCTreeSearch (mcts.py:132: 4 in the reference, behind its InferenceProxy): K
        simulations in flight per tree with virtual loss, a move takes iterations / K ticks (BatchedSelfPlay); 1 = sequential.
        eval_cache: the engine's per-slot evaluation cache (SelfPlayEngine(eval_cache=...)): positions a game slot has
        already sent to the same weights are not evaluated again -- the same games bit for bit in about half the network passes.
        On by default here (every epoch builds fresh engines after the weight refresh); False switches it off.
        trainer: "device" = the native SGD step (train.DeviceTrainer: TF32 / bf16 tensor-core convolutions, fp32 master weights,
        momentum buffers on the device), "torch" = PyTorch autograd with torch.optim.SGD, "auto" = "device" whenever the network
        is one it is built for and the replay memory is on the device.
        amp (trainer="torch" only): None = fp32 SGD steps; torch.float16 / torch.bfloat16 run forward/backward under autocast.  The reference's
        UpdateWorker trains under torch.cuda.amp.autocast() (fp16, updateworker.py:146-149) and keeps running 100-step rounds
        for as long as the epoch's self-play lasts; here exactly `updates_per_epoch` steps run after the epoch's self-play
        (deliberate: self-play is ~100x faster, so "as long as self-play runs" would mean almost no SGD steps per game) --
        scale `updates_per_epoch` to keep the reference's SGD-steps-per-game ratio.  The initial evaluate_policy(-1) of
        self_play_parallel.py:244 is available as train_model(initial_evaluation=True).
        lr_patience: ReduceLROnPlateau("max", patience, factor 0.5, min_lr 1e-5, cooldown 5) stepped with every epoch's
        evaluation reward (updateworker.py:66-68,96-98).
        deduplicate: UpdateWorker's option (updateworker.py:88-89): after every epoch's records are in, merge duplicate states.
        evaluation_opponent: None (evaluation_network, or the policy itself), "lookahead" or "random": the hard-coded
        evaluation_policy_container of the reference's train command (main.py:66, hardcoded_players.py)."""
        self.network, self.env, self.evaluation_network = network, env, evaluation_network
        self.iterations, self.epoch_length, self.initial_games, self.evaluation_games = iterations, epoch_length, initial_games, evaluation_games
        self.games_per_gpu, self.batch_size, self.updates_per_epoch = games_per_gpu, batch_size, updates_per_epoch
        self.alpha, self.seed, self.net, self.evaluation_opponent = alpha, seed, net, evaluation_opponent
        self.rank = torch.distributed.get_rank() if torch.distributed.is_initialized() else 0
        self.world = torch.distributed.get_world_size() if torch.distributed.is_initialized() else 1
        self.replay_kind, self.memory_step, self.max_memory_size = replay, memory_step, max_memory_size or memory_size
        self.memory = None if replay == "device" else Memory(memory_size)   # the device memory is created with the first games
        self._memory_size, self.deduplicate = memory_size, deduplicate
        # SGD(momentum 0.9, weight decay 1e-4): self_play_parallel.py:193
        self.optim = torch.optim.SGD(network.parameters(), lr=lr, momentum=momentum, weight_decay=weight_decay)
        self.lr_scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(self.optim, "max", patience=lr_patience, factor=0.5,
                                                                       min_lr=0.00001, cooldown=5)
        self.save_dir, self.save_memory_files, self._recent_memory_file, self.amp = save_dir, save_memory, None, amp
        from . import train as _train
        if trainer not in ("auto", "device", "torch"):
            raise ValueError("trainer must be 'auto', 'device' or 'torch'")
        can = replay == "device" and _train.supports(network)
        if trainer == "device" and not can:
            raise ValueError("trainer='device' needs replay='device' and a ResidualTower(7, 6, 7, filter_factor=32)")
        self.trainer_kind = "device" if (trainer == "device" or (trainer == "auto" and can)) else "torch"
        self._trainer, self._momentum, self._weight_decay = None, momentum, weight_decay
        self.search_threads = int(search_threads)
        self.eval_cache = eval_cache       # SelfPlayEngine(eval_cache=...): repeated positions are not re-evaluated (same games, fewer network passes)
        self.start_time = datetime.datetime.now().isoformat()                      # self_play_parallel.py:86
        self.games_played = 0
        self.history = []

    # ------------------------------------------------------------------ self-play on all ranks
    def _play(self, n_games_total, evaluate, update, generation):
        """n_games_total games (indices 0..n-1, swap_sides = index odd as in :250-253) sharded over the ranks."""
        G = min(self.games_per_gpu, max(2, -(-n_games_total // self.world)))
        G += G & 1
        sp = BatchedSelfPlay(self.network, env=self.env, n_games=G, sims=self.iterations, net=self.net,
                             evaluation_network=self.evaluation_network if evaluate else None, evaluate=evaluate, update=update,
                             alpha=self.alpha, seed=self.seed + 7919 * generation, rank=self.rank, world=self.world,
                             games_target=n_games_total, opponent=self.evaluation_opponent if evaluate else None,
                             search_threads=self.search_threads, eval_cache=self.eval_cache)
        device_replay = update and self.replay_kind == "device"
        if device_replay and self.memory is None:
            self.memory = DeviceReplay(sp.game, self._memory_size, self.max_memory_size, seed=self.seed)
        recs, res = [], []
        while True:
            sp.engine.run_ticks(sp.engine.safe_poll_interval)
            if device_replay:       # pull_from_queue without the queue: device ring -> device memory
                r = self.memory.drain_engine(sp.engine, append=self.world == 1)
                recs.append(r.clone() if self.world > 1 else r.shape[0])
            else:
                recs.append(sp.engine.drain_records())
            res.append(sp.engine.drain_results())
            if sp.engine.all_idle():
                break
        sp.engine.check_overflow()
        game = sp.game
        sp.close()
        dev = torch.device("cuda", torch.cuda.current_device())
        all_res = parallel.gather_structured(np.concatenate(res), dst=0, device=dev)         # replaces result_queue
        results = results_to_dicts(parallel.merge_results_in_game_order(all_res)) if self.rank == 0 else []
        if device_replay:
            if self.world == 1:
                return int(sum(recs)), results
            mine = torch.cat(recs) if recs else torch.empty(0, 80, dtype=torch.uint8, device=dev)
            everyone = parallel.gather_device_rows(mine, dst=0)                                # replaces memory_queue (NCCL)
            if self.rank != 0:
                return 0, []
            return int(sum(self.memory.append_records(t) for t in everyone)), results
        all_recs = parallel.gather_structured(np.concatenate(recs), dst=0, device=dev)       # replaces memory_queue
        if self.rank != 0:
            return [], []
        return records_to_moves(all_recs, game), results

    def _sync_weights(self):
        """Replaces the checkpoint-file hand-off (updateworker.py:111-117 -> inference_worker.py:68-73)."""
        if self.world == 1:
            return
        for p in list(self.network.parameters()) + list(self.network.buffers()):
            torch.distributed.broadcast(p.data, src=0)

    # ------------------------------------------------------------------ reference API
    def compare_models(self):
        """:355-379: epoch_length evaluation games policy vs evaluation policy, alternating sides."""
        _, results = self._play(self.epoch_length, evaluate=True, update=False, generation=10_000)
        return parse_results(results) if self.rank == 0 else (0, {})

    def evaluate_policy(self, epoch):
        """:329-353 (the evaluation-games half)."""
        _, results = self._play(self.evaluation_games, evaluate=True, update=False, generation=20_000 + epoch)
        return parse_results(results)[0] if self.rank == 0 else 0

    def update(self):
        """UpdateWorker.update (updateworker.py:141-149): `updates_per_epoch` SGD steps on uniform samples."""
        if self.rank != 0 or self.memory is None or len(self.memory) < self.batch_size:
            return None
        if self.trainer_kind == "device":
            from .train import DeviceTrainer
            if self._trainer is None:     # parameters, running statistics and momentum buffers stay on the device between epochs
                self._trainer = DeviceTrainer(self.network, batch_size=self.batch_size, lr=self.optim.param_groups[0]["lr"],
                                              momentum=self._momentum, weight_decay=self._weight_decay, seed=self.seed)
            loss = None
            for _ in range(self.updates_per_epoch):
                loss = self._trainer.step_from_batch(self.memory.sample_batch(self.batch_size), lr=self.optim.param_groups[0]["lr"])
            self._trainer.store(self.network)     # the module is what the evaluators pack and the checkpoints save
            return float(loss[0])
        self.network.train()
        last = None
        for _ in range(self.updates_per_epoch):
            with torch.autocast("cuda", dtype=self.amp or torch.bfloat16, enabled=self.amp is not None):
                if self.replay_kind == "device":
                    loss = loss_from_batch(self.network, self.memory.sample_batch(self.batch_size))
                else:
                    loss = mcts_loss(self.network, self.memory.sample(self.batch_size))
            self.optim.zero_grad()
            loss.backward()
            self.optim.step()
            last = float(loss.detach())
        self.network.eval()
        return last

    # ------------------------------------------------------------------ on-disk state (rank 0)
    def _host_memory(self):
        """The replay memory as the reference's picklable object: a Memory of Move tuples (CPU tensors)."""
        if self.replay_kind != "device":
            return self.memory
        m = Memory(self.memory.max_size)
        for mv in self.memory.to_moves():
            m.add(Move(*(t.cpu() for t in mv)))
        return m

    def save(self, games_played):
        """UpdateWorker's {"saved_name": ...} task: save_model + save_memory (updateworker.py:84-93,111-139)."""
        if self.rank != 0 or not self.save_dir:
            return None
        name = checkpoint.save_model(self.network, checkpoint.model_file_name(self.save_dir, self.start_time, games_played))
        if self.save_memory_files and self.memory is not None:
            self._recent_memory_file = checkpoint.save_memory(self._host_memory(), self.save_dir, self.start_time,
                                                              previous=self._recent_memory_file)
        return name

    def resume(self, resume_model=False, resume_memory=False):
        """BaseWorker.load_model / load_memory with prev_run=True (base_worker.py:26-42): newest files of the newest earlier run."""
        if not self.save_dir or not os.path.isdir(self.save_dir):
            return
        if self.rank == 0:
            try:
                if resume_memory:
                    old = checkpoint.load_memory(checkpoint.recent_save_file(self.save_dir, self.start_time, True, "memory"))
                    moves = list(getattr(old, "_buffer", old))
                    if self.replay_kind == "device":
                        if self.memory is None:
                            self.memory = DeviceReplay(game_id_of(self.env), self._memory_size, self.max_memory_size, seed=self.seed)
                        self.memory.extend(moves)
                    else:
                        for m in moves:
                            self.memory.add(m)
                if resume_model:
                    dev = next(self.network.parameters()).device
                    checkpoint.load_model(self.network, checkpoint.recent_save_file(self.save_dir, self.start_time, True, "model"), map_location=dev)
                    if self._trainer is not None:          # the device copies follow the module (a fresh optimizer, as in the reference)
                        self._trainer.load(self.network, reset_momentum=True)
            except (ValueError, OSError, ImportError, AttributeError) as e:
                # no earlier run (max() of an empty list), or an unreadable file: the reference logs the exception and goes on
                # (base_worker.py:31-42)
                import logging
                logging.getLogger(__name__).warning("resume: %s: %s", type(e).__name__, e)
        if resume_model:
            self._sync_weights()

    def train_model(self, num_epochs=10, resume_model=False, resume_memory=False, initial_evaluation=False):
        """:213-291: initial games, then per epoch: epoch_length self-play games -> update -> checkpoint -> weight sync ->
        evaluation -> LR schedule."""
        self.network.eval()
        self.resume(resume_model, resume_memory)
        gen = 0
        if initial_evaluation and self.evaluation_games:     # self_play_parallel.py:244
            self.history.append(dict(epoch=-1, evaluation_reward=self.evaluate_policy(-1)))
        self._remember(self._play(self.initial_games, evaluate=False, update=True, generation=gen)[0])
        import time

        def now():
            torch.cuda.synchronize()
            return time.time()
        for epoch in range(num_epochs):
            gen += 1
            t0 = now()
            moves, results = self._play(self.epoch_length, evaluate=False, update=True, generation=gen)
            self._remember(moves)
            t1 = now()
            if self.memory_step and self.rank == 0 and self.memory is not None:   # stagger_memory (updateworker.py:107-109)
                self.memory.change_size(min(self.memory.max_size + self.memory_step, self.max_memory_size))
            if self.deduplicate and self.rank == 0 and self.memory is not None:       # updateworker.py:88-89
                self.memory.deduplicate("state", ["actual_val", "tree_probs"], Move)
            self.games_played += self.epoch_length
            loss = self.update()
            t2 = now()
            saved = self.save(self.epoch_length * (epoch + 1))
            self._sync_weights()
            t3 = now()
            reward = self.evaluate_policy(epoch) if self.evaluation_games else 0
            if self.rank == 0:
                self.lr_scheduler.step(reward)                                         # updateworker.py:96-98
            t4 = now()
            self.history.append(dict(epoch=epoch, loss=loss, self_play=parse_results(results)[0] if self.rank == 0 else None,
                                     evaluation_reward=reward, memory=len(self.memory) if self.memory is not None else 0,
                                     saved_model=saved, lr=self.optim.param_groups[0]["lr"],
                                     seconds=dict(self_play=t1 - t0, update=t2 - t1, weight_sync=t3 - t2, evaluation=t4 - t3)))
        return self.history

    def _remember(self, moves):
        if self.replay_kind != "device":      # device replay: _play has already appended the records on the GPU
            for m in moves:
                self.memory.add(m)

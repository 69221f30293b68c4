"""Times the UNMODIFIED reference's own CPU self-play (staged by oracle/build_ref.py) -- the reference arm of bench.py.
TEST / BENCH INFRASTRUCTURE ONLY: nothing in the product package imports this.

Mode (i) of SURVEY.md 8(d), "direct": one process per host core, one torch thread each, every process looping
``SelfPlayer.play_episode(update=True)`` (selfplayworker.py:172-194) over two ``MCTreeSearch(network=ResidualTower-N,
env=Connect4Env, iterations=sims)`` (mcts.py:116-165) exactly as ``SelfPlayWorker.set_up_policies`` builds them
(selfplayworker.py:67-90).  A simulation = one ``search_node`` call (mcts.py:340-367), a position = one ``_play`` (mcts.py:272);
both are counted by thin subclass wrappers into shared memory, the run is time-boxed (a whole game takes ~9 minutes per core).
"""
import os
import sys
import time


import ctypes as _C


class Pair(_C.Structure):   # live counters of one worker (module level: a spawned child has to unpickle it)
    _fields_ = [("sims", _C.c_long), ("moves", _C.c_long)]


def _worker(idx, shared, blocks, sims, alpha):
    os.environ["CUDA_VISIBLE_DEVICES"] = ""        # mcts.py:18 picks cuda when available: this arm is the HOST-CPU reference
    import tempfile
    os.chdir(tempfile.mkdtemp(prefix="spx_ref_"))  # the reference's modules create log files / run folders in the CWD
    from oracle import build_ref
    for p in reversed(build_ref.import_paths()):
        if p not in sys.path:
            sys.path.insert(0, p)
    import numpy as np
    import torch
    torch.set_num_threads(1)
    import games.algos.mcts as mcts
    from games.algos.selfplayworker import SelfPlayer
    from games.connect4.connect4env import Connect4Env
    from games.general.modules import ResidualTower
    torch.manual_seed(0)
    np.random.seed(1000 + idx)
    net = ResidualTower(7, 6, 7, num_blocks=blocks).eval()

    class Sink:
        def put(self, *_a, **_k):
            pass

    class Counted(mcts.MCTreeSearch):
        def search_node(self):
            shared.sims += 1
            return super().search_node()

        def _play(self, *a, **k):
            shared.moves += 1
            return super()._play(*a, **k)
    trees = []
    for _ in (0, 1):
        t = Counted(network=net, env=Connect4Env, memory_queue=Sink(), iterations=sims, alpha=alpha)
        t.train(False)
        t.evaluate(False)
        trees.append(t)
    sp = SelfPlayer(trees[0], trees[1], Connect4Env(), Sink())
    g = idx
    with torch.no_grad():
        while True:
            sp.play_episode(swap_sides=bool(g & 1), update=True)
            g += 1


def time_reference(seconds, blocks, sims, procs=None, alpha=1.0):
    """`procs` processes (default: every host core) for `seconds`; returns sims/s and positions/s counted live."""
    import ctypes as C
    import multiprocessing as mp
    # spawn, not fork: the GPU arm of bench.py calls this with CUDA already initialised in the parent, and the reference picks its
    # device at import time (mcts.py:18) -- a forked child would inherit "cuda is available" and die in .to(device); a spawned
    # child starts clean and hides the GPUs before it imports torch
    ctx = mp.get_context("spawn")
    procs = procs or len(os.sched_getaffinity(0))

    shared = [ctx.RawValue(Pair) for _ in range(procs)]
    ps = [ctx.Process(target=_worker, args=(i, shared[i], blocks, sims, alpha), daemon=True) for i in range(procs)]
    for p in ps:
        p.start()
    t_dead = time.time() + 180
    while time.time() < t_dead and not all(s.sims > 0 for s in shared):
        if any(not p.is_alive() for p in ps):
            raise RuntimeError("a reference worker died during start-up")
        time.sleep(0.2)
    s0 = sum(s.sims for s in shared); m0 = sum(s.moves for s in shared); t0 = time.time()
    time.sleep(seconds)
    s1 = sum(s.sims for s in shared); m1 = sum(s.moves for s in shared); t1 = time.time()
    for p in ps:
        p.terminate()
    for p in ps:
        p.join(timeout=5)
    dt = t1 - t0
    return {"sims_per_s": (s1 - s0) / dt, "positions_per_s": (m1 - m0) / dt, "cores": procs, "seconds": dt, "sims": s1 - s0}


if __name__ == "__main__":
    import json
    print(json.dumps(time_reference(float(sys.argv[1]) if len(sys.argv) > 1 else 10.0, 20, 800,
                                    procs=int(sys.argv[2]) if len(sys.argv) > 2 else None)))

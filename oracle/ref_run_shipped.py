"""SURVEY.md 8(d) mode (ii), "as shipped": the UNMODIFIED reference's own multi-process self-play -- SelfPlayScheduler.compare_models
(self_play_parallel.py:355-379) with its SelfPlayWorker processes (8 games each, MCTreeSearch(thread_count=4) behind an
InferenceProxy) and ONE InferenceWorker batching the network on the CPU (inference_worker.py:89-119), spawn start method --
time-boxed; counted: requests the InferenceWorker answered (= network evaluations = simulations that needed the network).
TEST / BENCH INFRASTRUCTURE ONLY.   python -m oracle.ref_run_shipped [seconds] [blocks] [sims] [workers]
"""
import json
import os
import sys
import tempfile
import threading
import time


def _paths():
    here = os.path.dirname(os.path.abspath(__file__))
    root = os.path.dirname(here)
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import build_ref
    if build_ref.available():
        ps = build_ref.import_paths()
    else:
        ps = ["/root/reference", os.path.join(here, "_shims")]
    for p in reversed(ps):
        if p not in sys.path:
            sys.path.insert(0, p)


os.environ["CUDA_VISIBLE_DEVICES"] = ""       # the HOST-CPU reference (inference_proxy.py picks cuda when available)
_paths()                                       # module level: spawned children re-import this file first

import torch  # noqa: E402
from games.algos.inference_worker import InferenceWorker  # noqa: E402


class CountingInferenceWorker(InferenceWorker):
    """InferenceWorker whose request counter (inference_worker.py:112) is mirrored into shared memory."""
    shared = None

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        self.shared = CountingInferenceWorker.shared

    def distribute(self, queues, evaluator):
        before = self.counter
        super().distribute(queues, evaluator)
        if self.counter != before:
            self.shared.value = self.counter


def main(seconds=30.0, blocks=20, sims=800, workers=None):
    os.chdir(tempfile.mkdtemp(prefix="spx_ref_shipped_"))
    torch.multiprocessing.set_start_method("spawn", force=True)
    import games.algos.self_play_parallel as spp
    from games.algos.mcts import MCTreeSearch
    from games.connect4.connect4env import Connect4Env
    from games.general.base_model import ModelContainer
    from games.general.modules import ResidualTower
    workers = workers or os.cpu_count()
    CountingInferenceWorker.shared = torch.multiprocessing.Value("l", 0)
    spp.InferenceWorker = CountingInferenceWorker
    torch.manual_seed(0)
    net = ResidualTower(7, 6, 7, num_blocks=blocks)
    torch.manual_seed(1)
    net2 = ResidualTower(7, 6, 7, num_blocks=blocks)
    kw = dict(env=Connect4Env, iterations=sims, thread_count=4)
    os.mkdir("saves")
    sched = spp.SelfPlayScheduler(ModelContainer(MCTreeSearch, policy_kwargs=dict(kw)), Connect4Env,
                                  evaluation_policy_container=ModelContainer(MCTreeSearch, policy_kwargs=dict(kw)), network=net,
                                  evaluation_network=net2, save_dir="saves", epoch_length=100000)
    th = threading.Thread(target=lambda: sched.compare_models(num_workers=max(3, workers), threads_per_worker=8), daemon=True)
    th.start()
    c = CountingInferenceWorker.shared
    t_dead = time.time() + 240
    while c.value == 0 and time.time() < t_dead:
        time.sleep(0.5)
    if c.value == 0:
        print(json.dumps({"error": "no request was answered within 240 s"}), flush=True)
        os._exit(1)
    time.sleep(5.0)                                  # past the start-up transient
    c0, t0 = c.value, time.time()
    time.sleep(seconds)
    c1, t1 = c.value, time.time()
    print(json.dumps({"leaf_evals_per_s": (c1 - c0) / (t1 - t0), "seconds": t1 - t0, "processes": max(3, workers), "threads_per_worker": 8,
                      "thread_count": 4, "evals": c1 - c0}), flush=True)
    import multiprocessing as mp
    for p in mp.active_children():
        p.terminate()
    os._exit(0)


if __name__ == "__main__":
    a = sys.argv[1:]
    main(float(a[0]) if len(a) > 0 else 30.0, int(a[1]) if len(a) > 1 else 20, int(a[2]) if len(a) > 2 else 800, int(a[3]) if len(a) > 3 else None)

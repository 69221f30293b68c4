#!/usr/bin/env python
"""bench.py -- MCTS sims/sec (and self-play positions/sec) for Connect4 self-play, 800 sims/move.

Contract (see DESIGN.md "Measurement"):
  python bench.py --gpus N --steps K --warmup W            (N>1: launched under torch.distributed.run)
  python bench.py --impl reference ...                     (the CPU arm: oracle port on the host cores)

Workload at N=1 = BASELINE.json configs[1]: Connect4 6x7 self-play, the repo's connect4 net
ResidualTower(7,6,7,num_blocks=20) with random init (torch.manual_seed(0)), 1024 concurrent games per GPU,
800 sims/move, Dirichlet noise (--alpha, default 1) generated on device, tie noise on, finished games replaced at once.
A "step" = 800 engine ticks (one tick = one spx_advance over all games + one batched network evaluation),
i.e. about one searched move per game.  `value` = completed MCTS simulations per second over all ranks, device
timed with everything resident in HBM.  `e2e` = the same metric through the package's public API with HOST
buffers: every step uploads the packed network weights from pinned host memory (the weight refresh the
reference does through checkpoint files) and downloads that step's Move records, game results and counters.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

FLOP_PER_LEAF_20 = 497_138_048  # ResidualTower-20, SURVEY.md 6 / 8(d)
TICKS_PER_STEP = 800


def flop_per_leaf(blocks, conv_only=False):
    conv = 2 * 42 * (9 * 3 * 128 + blocks * 2 * 9 * 128 * 128 + 128 * 64)
    fc = 2 * (1344 * 7 + 1344 * 256 + 256)
    return conv if conv_only else conv + fc


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--games", type=int, default=None, help="concurrent games per GPU (default: 1024 on one GPU = BASELINE configs[1]; "
                    "16384 / N on N > 1 GPUs = configs[2])")
    ap.add_argument("--no-stagger", action="store_true", help="skip the quick-games phase that decorrelates the slots' game phases before warm-up")
    ap.add_argument("--no-config4", action="store_true", help="skip the configs[3] leg (4096 head-to-head games, 400 sims/move, two towers)")
    ap.add_argument("--no-config5", action="store_true", help="skip the configs[4] leg (one epoch of the full loop: self-play, SGD, weight refresh, evaluation)")
    ap.add_argument("--sims", type=int, default=800)
    ap.add_argument("--blocks", type=int, default=20)
    ap.add_argument("--net", default="tower", choices=["tower", "torch", "hash"])
    ap.add_argument("--ticks-per-step", type=int, default=TICKS_PER_STEP)
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="bounded CPU-baseline sample per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-aux-rooflines", action="store_true", help="skip the search-only and env-kernel roofline legs (tests)")
    ap.add_argument("--max-sims-per-tick", type=int, default=8)
    ap.add_argument("--no-fused", action="store_true", help="launch spx_advance + the network kernel per tick instead of the fused tick kernel")
    ap.add_argument("--fused-chunk", type=int, default=400, help="ticks per launch of the fused tick kernel")
    ap.add_argument("--alpha", type=float, default=1.0, help="Dirichlet alpha of the root noise (mcts.py:135 default 1; tictactoeconfig.py:9 uses 0.15)")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            try:
                pw.append(float(f[2]))
            except ValueError:
                pass
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort(); pw.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w": pw[len(pw) // 2] if pw else None}


# ------------------------------------------------------------------------------------------------- CPU arm
def _cpu_worker(idx, shared, blocks, sims, seed, alpha=1.0):
    """One host core: the oracle port (C tree/env restatement + fp32 torch net, batch 1, one thread), i.e. the
    reference's direct mode (BASELINE.md 4 mode (i)): SelfPlayer.play_episode over two MCTreeSearch."""
    import ctypes as C
    import numpy as np
    import torch
    torch.set_num_threads(1)
    from oracle import oracle as ox
    from self_play_reinforcement_learning_b200 import nets
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=blocks).eval()
    ox.lib().ox_set_live_counters(C.addressof(shared))   # the C oracle bumps (sims, moves) in this shared pair as it plays

    def fn(state, tree):
        with torch.no_grad():
            p, v = net.forward(torch.from_numpy(np.ascontiguousarray(state).astype(np.int64)))
        return p[0].tolist(), float(v.item())
    pynet = ox.PyNet(0, fn)
    g = idx
    while True:
        cfg = ox.make_cfg(0, sims, seed=seed, game_uid=g, noise_mode=2, alpha=alpha)
        ox.play_episode(cfg, bool(g & 1), nets=((pynet.addr, None), (pynet.addr, None)))
        g += 1 << 20


def reference_kind():
    """"reference": the unmodified reference staged as byte-code under oracle/_ref (oracle/build_ref.py) is importable;
    "port": only the C/torch restatement is (the fallback the sample string then names)."""
    from oracle import build_ref
    return "reference" if build_ref.available() and os.environ.get("SPX_BENCH_CPU_ARM", "") != "port" else "port"


def cpu_baseline(seconds, blocks, sims, procs=None, alpha=1.0, kind=None):
    """Time-boxed: `procs` processes play self-play games for `seconds`; sims and moves counted live.
    kind "reference": oracle/ref_run.py (the reference's own MCTreeSearch / SelfPlayer / ResidualTower); "port": the oracle."""
    if (kind or reference_kind()) == "reference":
        from oracle import ref_run
        return dict(ref_run.time_reference(seconds, blocks, sims, procs=procs, alpha=alpha), kind="reference")
    return dict(_cpu_baseline_port(seconds, blocks, sims, procs=procs, alpha=alpha), kind="port")


def cpu_baseline_as_shipped(seconds, blocks, sims):
    """SURVEY.md 8(d) mode (ii): the unmodified reference's own multi-process path (SelfPlayScheduler.compare_models: SelfPlayWorker
    processes with 8 threaded games each + one InferenceWorker batching the network on the CPU, spawn) -- oracle/ref_run_shipped.py in
    a subprocess with a hard time limit; network evaluations per second.  None when the staged reference is absent; an error
    string instead of a number when the run fails (the reference swallows its workers' exceptions)."""
    if reference_kind() != "reference":
        return None
    try:
        import signal
        # its own session: whatever happens, the whole tree of reference worker processes goes away with it
        p = subprocess.Popen([sys.executable, "-m", "oracle.ref_run_shipped", str(seconds), str(blocks), str(sims)], cwd=ROOT,
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, start_new_session=True)
        try:
            out, _ = p.communicate(timeout=420)
        finally:
            try:
                os.killpg(p.pid, signal.SIGKILL)
            except OSError:
                pass
        lines = [ln for ln in out.splitlines() if ln.startswith("{")]
        d = json.loads(lines[-1])
        if "error" in d:
            return {"error": d["error"]}
        return {"leaf_evals_per_s": d["leaf_evals_per_s"], "unit": "network evaluations/s", "processes": d["processes"], "threads_per_worker": d["threads_per_worker"],
                "thread_count": d["thread_count"], "seconds": d["seconds"],
                "what": "SelfPlayScheduler.compare_models as shipped (self_play_parallel.py:355-379): SelfPlayWorker processes + one InferenceWorker "
                        "on the CPU, requests answered by the InferenceWorker per second (inference_worker.py:112)"}
    except Exception as e:      # noqa: BLE001 -- a reported baseline must not take the bench down
        return {"error": f"{type(e).__name__}: {e}"[:300]}


def _sample_text(r, blocks, sims):
    who = ("the UNMODIFIED reference (games/algos/mcts.py MCTreeSearch x 2 + selfplayworker.SelfPlayer.play_episode + "
           f"games/general/modules.py ResidualTower-{blocks} in fp32 at batch 1, byte-compiled into oracle/_ref)") if r["kind"] == "reference" else \
          (f"the oracle PORT (C tree/env restatement + fp32 torch ResidualTower-{blocks} at batch 1; the staged reference oracle/_ref is absent)")
    return (f"{r['seconds']:.0f}s window, {r['cores']} processes x 1 torch thread, CUDA hidden, each looping Connect4 self-play episodes "
            f"(two trees, {sims} sims/move) with {who}: the reference's direct mode (SURVEY.md 8d mode i)")


def _cpu_baseline_port(seconds, blocks, sims, procs=None, alpha=1.0):
    import multiprocessing as mp
    import ctypes as C
    ctx = mp.get_context("fork")
    procs = procs or len(os.sched_getaffinity(0))

    class Pair(C.Structure):
        _fields_ = [("sims", C.c_long), ("moves", C.c_long)]
    shared = [ctx.RawValue(Pair) for _ in range(procs)]
    ps = [ctx.Process(target=_cpu_worker, args=(i, shared[i], blocks, sims, 0, alpha), daemon=True) for i in range(procs)]
    for p in ps:
        p.start()
    # wait until every worker is past start-up (first simulation done), then measure a clean window
    t_dead = time.time() + 120
    while time.time() < t_dead and not all(s.sims > 0 for s in shared):
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


def run_reference(args, workload, out=sys.stdout):
    """--impl reference: the reference's own CPU implementation of the path on all host cores (kind "reference": the unmodified
    reference byte-compiled into oracle/_ref; "port": the oracle restatement, only when oracle/_ref is absent); each 'step' is
    a bounded time-boxed sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals, cores = [], None
    # a "step" is a bounded time-boxed sample; the whole run stays within ~2 minutes whatever K is (>= 2 s, <= --cpu-seconds each)
    per_step = max(2.0, min(args.cpu_seconds, 90.0 / max(args.steps, 1)))
    for _ in range(args.warmup if args.warmup < 1 else 1):
        cpu_baseline(min(per_step, 5.0), args.blocks, args.sims, alpha=args.alpha)
    t0 = time.time()
    for _ in range(args.steps):
        r = cpu_baseline(per_step, args.blocks, args.sims, alpha=args.alpha)
        vals.append(r)
        cores = r["cores"]
    tot_sims = sum(v["sims"] for v in vals); tot_t = sum(v["seconds"] for v in vals)
    value = tot_sims / tot_t
    sample = f"{args.steps} windows: " + _sample_text(vals[-1], args.blocks, args.sims)
    line = {"impl": "reference", "metric": "mcts_sims_per_sec", "value": value, "unit": "sims/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * tot_t / max(args.steps, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload,
            "cpu_baseline": {"value": value, "unit": "sims/s", "cores": cores, "kind": vals[-1]["kind"], "sample": sample},
            "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "positions_per_sec": sum(v["positions_per_s"] * v["seconds"] for v in vals) / tot_t,
            "wall_s": time.time() - t0}
    line["cpu_baseline"]["as_shipped"] = cpu_baseline_as_shipped(min(per_step, 20.0), args.blocks, args.sims)
    print(json.dumps(line), file=out, flush=True)


# ------------------------------------------------------------------------------------------------- GPU arm
def _claim_stdout():
    """The contract is ONE JSON line on stdout: libraries that write to fd 1 (NCCL prints its version banner there under
    NCCL_DEBUG=VERSION) are redirected to stderr; the returned file object is the real stdout for the final line."""
    sys.stdout.flush()
    real = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    return real


def main():
    args = parse()
    json_out = _claim_stdout()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    n_gpus = max(world, args.gpus)
    fixed_total = args.games is None and n_gpus > 1
    if args.games is None:   # configs[1] on one GPU; configs[2] (16384 concurrent games in total) sharded over N > 1 GPUs
        args.games = 1024 if n_gpus == 1 else 16384 // n_gpus
    workload = {"workload": f"connect4 6x7 self-play, ResidualTower-{args.blocks} random init (seed 0), {args.games} concurrent games/GPU"
                            + (f" ({args.games * n_gpus} in total = BASELINE configs[2])" if n_gpus > 1 else " (BASELINE configs[1])") +
                            f", {args.sims} sims/move, Dirichlet alpha {args.alpha:g}, two trees per game, finished games replaced immediately, "
                            "game phases of the slots decorrelated before warm-up",
                "games_per_gpu": args.games, "sims_per_move": args.sims, "net": args.net, "ticks_per_step": args.ticks_per_step,
                "l2_policy": f"node pools ({args.games * 2 * 2.7e-3:.1f} GB/GPU) and weights (12.6 MB) are the inputs; the touched working set per step "
                             "exceeds L2 (126 MB), no flush needed",
                "parallelism": f"games sharded over {n_gpus} GPU(s), no data-path collective" +
                               ("; e2e: rank 0 uploads the weights and ncclBroadcast()s them, records are gathered to rank 0 over NCCL every step" if n_gpus > 1 else "")}
    if args.impl == "reference":
        return run_reference(args, workload, json_out)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from self_play_reinforcement_learning_b200 import _lib, nets
    from self_play_reinforcement_learning_b200.engine import HashNetEvaluator, SelfPlayEngine
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay

    torch.manual_seed(0)
    module = nets.ResidualTower(7, 6, 7, num_blocks=args.blocks).eval()
    G, T = args.games, args.ticks_per_step

    # ---- public-API object (also used for the resident-data measurement through its engine)
    sp = BatchedSelfPlay(module, game=0, n_games=G, sims=args.sims, net=args.net, seed=0, rank=rank, world=world,
                         max_sims_per_tick=args.max_sims_per_tick, alpha=args.alpha)
    eng, ev = sp.engine, sp.evaluator
    blob_host = sp.packed_weights_pinned() if args.net == "tower" else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    fused = args.net == "tower" and not args.no_fused and getattr(ev, "fused_ticks", None) is not None \
        and ev.tower.ncta == 2 and ev.tower.fused_heads and os.environ.get("SPX_FUSED_TICK", "1") != "0"
    # ---- decorrelate the slots (untimed): a few generations of quick games, so that the timed window sees the steady-state mix
    # of game phases instead of "all games at ply k" (sims/s drifts with the phase: late plies re-visit terminal nodes for free)
    if not args.no_stagger:
        eng.stagger()
    # ---- warm-up (untimed)
    for _ in range(max(args.warmup, 3)):
        eng.run_ticks(T, fused=fused)
    barrier()
    c0 = eng.counters()
    launches0 = _lib.lib().spx_launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()

    # ---- timed region 1: everything resident (value).  Default: the fused tick kernel, `--fused-chunk` ticks per launch, CUDA
    # events around every launch (the dominant kernel IS the step).  --no-fused / other nets: one advance + one network launch
    # per tick, events around the kernels of every 8th tick.
    tower_ev, adv_ev, fused_ev = [], [], []
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    start.record()
    for s in range(args.steps):
        if fused:
            done = 0
            while done < T:
                n = min(args.fused_chunk, T - done)
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                eng.run_ticks(n, fused=True, chunk=n)
                f1.record()
                fused_ev.append((f0, f1, n))
                done += n
            continue
        for t in range(T):
            if args.net == "tower" and t % 8 == 0:
                a = (_lib.Event(), _lib.Event())
                n3 = tuple(_lib.Event() for _ in range(3))
                eng.tick(advance_events=a, net_events=n3)
                adv_ev.append(a); tower_ev.append(n3)
            else:
                eng.tick()
    end.record()
    barrier()
    ms = start.elapsed_time(end)
    launches1 = _lib.lib().spx_launch_count()   # kernels launched inside the timed region (the counters read below is outside)
    c1 = eng.counters()
    clocks = sampler.stop() if rank == 0 else None
    sims = c1["sims"] - c0["sims"]
    moves = c1["moves"] - c0["moves"]
    evals = c1["leaf_evals"] - c0["leaf_evals"]
    ticks = c1["ticks"] - c0["ticks"]
    path = (c1["path_len_sum"] - c0["path_len_sum"]) / max(sims, 1)
    stat = torch.tensor([ms, float(sims), float(moves), float(evals)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = stat.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = stat.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms, sims_all, moves_all, evals_all = mx[0].item(), sm[1].item(), sm[2].item(), sm[3].item()
    else:
        sims_all, moves_all, evals_all = float(sims), float(moves), float(evals)
    value = sims_all / (ms / 1e3)

    tower_ms = heads_ms = adv_ms = fused_tick_ms = None
    if fused_ev:
        fused_tick_ms = sum(f0.elapsed_time(f1) for f0, f1, _ in fused_ev) / sum(n for _, _, n in fused_ev)   # kernel time per tick
        # the two halves of a tick separately (explains the fused number; outside the timed region): 64 unfused, instrumented ticks
        for _ in range(64):
            a = (_lib.Event(), _lib.Event())
            n3 = tuple(_lib.Event() for _ in range(3))
            eng.tick(advance_events=a, net_events=n3)
            adv_ev.append(a); tower_ev.append(n3)
        torch.cuda.synchronize()
    if tower_ev:
        tower_ms = sum(e[0].elapsed_time(e[1]) for e in tower_ev) / len(tower_ev)
        heads_ms = sum(e[1].elapsed_time(e[2]) for e in tower_ev) / len(tower_ev)
        adv_ms = sum(a[0].elapsed_time(a[1]) for a in adv_ev) / len(adv_ev)

    # ---- timed region 2: end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        barrier()
        t_sims0 = eng.counters()["sims"]
        h2d = d2h = 0
        start2, end2 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        coll_ev, gathered, first_host = [], 0, None
        if world > 1:
            from self_play_reinforcement_learning_b200 import parallel
            from self_play_reinforcement_learning_b200.replay import DeviceReplay
            blob_dev = torch.empty(blob_host.numel(), dtype=torch.uint8, device="cuda")
            stagebuf = DeviceReplay(0, 16, 16, seed=0)        # only its device staging area is used (drain, no append)
            dist.broadcast(blob_dev, src=0)                   # untimed: NCCL sets up its channels on the first call of each kind
            parallel.gather_device_rows(torch.zeros(8, 80, dtype=torch.uint8, device="cuda"), dst=0)
            barrier()
        start2.record()
        for s in range(args.steps):
            if world == 1:
                out = sp.play_step(T, weights_host=blob_host, fused=fused)   # H2D weights, T ticks, D2H records/results/counters
                h2d += out["h2d_bytes"]; d2h += out["d2h_bytes"]
                continue
            # N > 1 (north_star): rank 0 uploads the new weights once and ncclBroadcast()s them over NVLink; after the ticks every
            # rank's records go to rank 0 over NCCL (device to device) and leave the box from there
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            if rank == 0:
                blob_dev.copy_(blob_host, non_blocking=True); h2d += blob_host.numel()
            e[0].record()
            dist.broadcast(blob_dev, src=0)
            e[1].record()
            sp.load_weights(blob_dev)
            eng.run_ticks(T, fused=fused)
            rows = stagebuf.drain_engine(eng, append=False)
            res = eng.drain_results(); cnt = eng.counters()
            d2h += res.nbytes + 80 + 16
            e[2].record()
            parts = parallel.gather_device_rows(rows, dst=0)
            e[3].record()
            if rank == 0:
                host = [p.cpu() for p in parts]
                gathered += sum(int(h.shape[0]) for h in host)
                d2h += sum(h.numel() for h in host)
                if s == 0:
                    first_host = host          # checked after the timed region
            coll_ev.append(e)
        end2.record()
        barrier()
        ms2 = start2.elapsed_time(end2)
        if world > 1 and rank == 0:
            _check_gathered(first_host, world, G)
        s2 = eng.counters()["sims"] - t_sims0
        st2 = torch.tensor([ms2, float(s2)], dtype=torch.float64, device="cuda")
        if world > 1:
            mx2 = st2.clone(); dist.all_reduce(mx2, op=dist.ReduceOp.MAX)
            sm2 = st2.clone(); dist.all_reduce(sm2, op=dist.ReduceOp.SUM)
            ms2, s2 = mx2[0].item(), sm2[1].item()
        e2e = {"value": s2 / (ms2 / 1e3), "unit": "sims/s", "h2d_bytes_per_step": h2d // args.steps, "d2h_bytes_per_step": d2h // args.steps,
               "ms_per_step": ms2 / args.steps}
        if world > 1:
            bc = sum(e[0].elapsed_time(e[1]) for e in coll_ev) / len(coll_ev)
            ga = sum(e[2].elapsed_time(e[3]) for e in coll_ev) / len(coll_ev)
            e2e["collectives_ms"] = {"weight_broadcast": bc, "record_gather": ga, "per": "step", "blob_bytes": int(blob_host.numel()),
                                     "records_gathered_per_step": gathered / args.steps}
            e2e["collectives_checked"] = _check_collectives(dist, parallel, blob_dev, rank, world, local_rank)

    # ---- search-only legs (hash net): HBM roofline of the search kernel at the workload's G and with 16x more trees
    search = None
    if rank == 0 and args.net == "tower" and not args.no_aux_rooflines:
        sp.close()
        peaks_s = _peaks()

        def search_leg(n_games, ticks):
            e2 = SelfPlayEngine(game=0, n_games=n_games, sims=args.sims, evaluator=HashNetEvaluator(0, 1), seed=1, noise_mode=2)
            e2.run_ticks(3 * ticks // 2)
            torch.cuda.synchronize()
            k0 = e2.counters()
            evs = []
            for t in range(ticks):
                a = (_lib.Event(), _lib.Event())
                e2.tick(advance_events=a)
                evs.append(a)
            torch.cuda.synchronize()
            k1 = e2.counters()
            a_ms = sum(a[0].elapsed_time(a[1]) for a in evs)
            s_ = k1["sims"] - k0["sims"]
            L = (k1["path_len_sum"] - k0["path_len_sum"]) / max(s_, 1)
            bytes_per_sim = 136.0 * L + 212.0
            gbs = s_ * bytes_per_sim / (a_ms / 1e3) / 1e9
            e2.close()
            return {"games": n_games, "sims_per_s_kernel_only": s_ / (a_ms / 1e3), "mean_path_len": L, "bytes_per_sim": bytes_per_sim,
                    "achieved_GBps": gbs, "peak_GBps": peaks_s["hbm_gbs"], "frac": gbs / peaks_s["hbm_gbs"],
                    "advance_ms_per_launch": a_ms / len(evs)}
        search = search_leg(G, 2 * T)
        free_b, _tot = torch.cuda.mem_get_info()
        if free_b > 110e9:   # 16384 games x 2 trees x 2.7 MB node pool = 88 GB
            search["with_16384_games"] = search_leg(16384, 600)

    env_roof = None
    if rank == 0 and args.net == "tower" and not args.no_aux_rooflines:
        env_roof = _env_roofline(_lib, torch)
    config4 = None
    if rank == 0 and world == 1 and args.net == "tower" and not args.no_aux_rooflines and not args.no_config4:
        sp.close()
        config4 = _config4_leg(torch, nets, BatchedSelfPlay, args.blocks)
    cache_leg = None
    if rank == 0 and world == 1 and args.net == "tower" and fused and not args.no_aux_rooflines and not args.no_config4:
        cache_leg = _eval_cache_leg(torch, nets, BatchedSelfPlay, args.blocks, args.games, args.sims, args.fused_chunk, value)
    train_leg = None
    if rank == 0 and world == 1 and args.net == "tower" and not args.no_aux_rooflines and not args.no_config4:
        train_leg = _train_leg(torch, nets, args.blocks)
    config5 = None
    if rank == 0 and world == 1 and args.net == "tower" and not args.no_aux_rooflines and not args.no_config5:
        config5 = _config5_leg(torch, nets, args.blocks, args.sims)
    if rank == 0:
        peaks = _peaks()
        tw = getattr(ev, "tower", None)
        fused_heads = bool(getattr(tw, "fused_heads", False))
        fl = flop_per_leaf(args.blocks, conv_only=not fused_heads)   # with fused heads the tower kernel is the whole network
        roof = None
        if tower_ms:
            leaves_per_launch = evals / max(ticks, 1)
            name = ("spx::tower::tower_kernel<2> (tcgen05 cta_group::2, SM pair" + (", fused FC heads)" if fused_heads else ")")) if getattr(tw, "ncta", 2) == 2 else "spx::tower::tower_kernel<1>"
            k_ms = tower_ms
            if fused_tick_ms:   # the step is ONE kernel: network + search engine; its time per tick is the denominator
                name = "spx::tower::tower_kernel<2, connect4, ENGINE> (fused tick: tcgen05 network with fused FC heads + the per-game search state machine in the same persistent CTAs)"
                k_ms = fused_tick_ms
            achieved = leaves_per_launch * fl / (k_ms / 1e3) / 1e12
            roof = {"bound": "tensor", "kernel": name, "achieved": achieved, "peak": peaks["bf16_tflops_sustained"],
                    "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops_sustained"], "traffic": _ncu_traffic(bool(fused_tick_ms)),
                    "peak_source": peaks["source"], "flop_per_leaf": fl, "leaves_per_launch": leaves_per_launch,
                    "per": "tick (a fused launch runs --fused-chunk ticks; time, leaves and traffic are per tick)" if fused_tick_ms else "launch",
                    "kernel_ms": k_ms, "heads_kernel_ms": None if fused_heads else heads_ms,
                    "tower_kernel_ms_unfused": tower_ms, "advance_kernel_ms_unfused": adv_ms,
                    "kernel_share_of_step": (k_ms * ticks / max(ms, 1e-9)) if world == 1 else None}
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            r = cpu_baseline(args.cpu_seconds, args.blocks, args.sims, alpha=args.alpha)
            cpu = {"value": r["sims_per_s"], "unit": "sims/s", "cores": r["cores"], "kind": r["kind"],
                   "sample": _sample_text(r, args.blocks, args.sims), "positions_per_s": r["positions_per_s"],
                   "as_shipped": cpu_baseline_as_shipped(min(args.cpu_seconds, 20.0), args.blocks, args.sims) if not args.no_aux_rooflines else None}
        line = {"metric": "mcts_sims_per_sec", "value": value, "unit": "sims/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if fixed_total else "weak",   # N > 1 default: 16384 concurrent games in total at every N (configs[2])
                "vs_baseline": None,
                "dtype": ("f16" if getattr(getattr(ev, "tower", None), "f16", False) else "bf16") if args.net == "tower" else ("bf16" if args.net == "torch" else "f64"),
                "data": "synthetic", "config": workload, "positions_per_sec": moves_all / (ms / 1e3), "leaf_evals_per_sec": evals_all / (ms / 1e3),
                "mean_select_path_len": path, "fused_tick_kernel": bool(fused_tick_ms), "fused_launches": (f"work-conserving (spx_tick_fused_balanced), {args.fused_chunk} ticks per launch: a step is ticks_per_step x ceil(games / 14) network passes, a game gets ticks_per_step ticks on average" if fused_tick_ms and os.environ.get("SPX_TICK_BALANCE", "1") != "0" else None), "e2e": e2e, "gpu_launches": int(launches1 - launches0), "clocks": clocks,
                "roofline": roof, "search_roofline": search, "env_roofline": env_roof, "config4_head_to_head": config4, "eval_cache": cache_leg, "train_step": train_leg, "config5_epoch": config5, "cpu_baseline": cpu}
        print(json.dumps(line), file=json_out, flush=True)
    if world > 1:
        dist.destroy_process_group()


def _check_gathered(host_parts, world, games_per_rank):
    """Every gathered record must come from the rank that owns its game (slot sharding, parallel.owner_of_game) and be a
    well-formed Move: visit distribution sums to 1, |actual_val| <= 1, piece counts of the two sides differ by at most one."""
    import numpy as np
    from self_play_reinforcement_learning_b200.engine import RECORD_DTYPE
    from self_play_reinforcement_learning_b200.parallel import owner_of_game
    for r, part in enumerate(host_parts):
        rec = np.frombuffer(part.numpy().tobytes(), dtype=RECORD_DTYPE)
        if len(rec) == 0:
            continue
        own = owner_of_game(rec["game_index"].astype(np.int64), world, games_per_rank)
        if not (own == r).all():
            raise SystemExit(f"bench: records gathered from rank {r} belong to other ranks' games")
        pop = lambda x: np.unpackbits(np.ascontiguousarray(x).view(np.uint8).reshape(len(x), 8), axis=1).sum(1).astype(np.int64)   # noqa: E731
        if not (np.abs(rec["tree_probs"].sum(1) - 1.0) < 1e-4).all() or not (np.abs(rec["actual_val"]) <= 1).all() \
                or not (np.abs(pop(rec["own"]) - pop(rec["opp"])) <= 1).all() or (rec["own"] & rec["opp"]).any():
            raise SystemExit(f"bench: malformed record in the gather from rank {r}")


def _check_collectives(dist, parallel, blob_dev, rank, world, local_rank):
    """Outside the timed region, N > 1 only (the driver's 1-GPU box skips the 2-GPU pytest): (1) every rank holds rank 0's
    weight blob after the broadcast (64-bit checksums all-gathered and compared); (2) records of 8 hash-net games played to
    the end on every rank with globally sharded game indices, gathered over NCCL, are byte-identical to the same games
    replayed locally on rank 0 (the engine itself is held against the oracle by tests/)."""
    import numpy as np
    import torch
    from self_play_reinforcement_learning_b200.engine import HashNetEvaluator, SelfPlayEngine, RECORD_DTYPE
    from self_play_reinforcement_learning_b200.replay import DeviceReplay
    n8 = blob_dev.numel() // 8
    words = blob_dev[:n8 * 8].view(torch.int64)
    ck = (words * (torch.arange(n8, device="cuda", dtype=torch.int64) % 251 + 1)).sum().reshape(1)   # wrapping int64 arithmetic
    cks = [torch.zeros_like(ck) for _ in range(world)]
    dist.all_gather(cks, ck)
    if any(int(c.item()) != int(cks[0].item()) for c in cks):
        raise SystemExit("bench: weight blob differs between ranks after ncclBroadcast")

    def play(offset):
        e = SelfPlayEngine(game=0, n_games=8, sims=40, evaluator=HashNetEvaluator(0, 1), seed=11, noise_mode=2,
                           slot_offset=offset, slot_stride=8 * world, games_target=8 * world)
        stage = DeviceReplay(0, 16, 16, seed=0)
        rows = []
        while True:
            e.run_ticks(64)
            rows.append(stage.drain_engine(e, append=False).clone())
            e.drain_results()
            if e.all_idle():
                break
        e.close()
        return torch.cat(rows)
    parts = parallel.gather_device_rows(play(8 * rank), dst=0)
    if rank == 0:
        key = lambda a: np.sort(np.frombuffer(a.cpu().numpy().tobytes(), dtype=RECORD_DTYPE), order=["game_index", "tree", "ply"]).tobytes()   # noqa: E731
        for r in range(1, world):
            if key(parts[r]) != key(play(8 * r)):
                raise SystemExit(f"bench: records gathered from rank {r} differ from the same games replayed on rank 0")
    return {"broadcast_checksum_equal_on_all_ranks": True, "gathered_records_equal_local_replay": True, "ranks": world}


def _config4_leg(torch, nets, BatchedSelfPlay, blocks, games=4096, sims=400, ticks=600, eval_cache=0):
    """BASELINE configs[3]: elo.py head-to-head evaluation, two random-init ResidualTower nets, 4096 games, 400 sims/move
    (evaluate mode, no records), both towers native; leaves are partitioned by owning network on the device."""
    torch.manual_seed(0)
    a = nets.ResidualTower(7, 6, 7, num_blocks=blocks).eval()
    torch.manual_seed(1)
    b = nets.ResidualTower(7, 6, 7, num_blocks=blocks).eval()
    sp = BatchedSelfPlay(a, game=0, n_games=games, sims=sims, net="tower", evaluation_network=b, evaluate=True, update=False, seed=0, eval_cache=eval_cache)
    sp.engine.stagger()
    sp.engine.run_ticks(sims)
    torch.cuda.synchronize()
    c0 = sp.engine.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sp.engine.run_ticks(ticks)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c1 = sp.engine.counters()
    out = {"workload": f"{games} head-to-head games, {sims} sims/move, evaluate mode, two native ResidualTower-{blocks}", "ticks": ticks,
           "sims_per_s": (c1["sims"] - c0["sims"]) / (ms / 1e3), "moves_per_s": (c1["moves"] - c0["moves"]) / (ms / 1e3),
           "games_per_s": (c1["games_finished"] - c0["games_finished"]) / (ms / 1e3), "ms_per_tick": ms / ticks,
           "leaf_evals_per_tick": (c1["leaf_evals"] - c0["leaf_evals"]) / ticks}
    sp.close()
    if eval_cache:
        out["hit_rate"] = (c1["cache_hits"] - c0["cache_hits"]) / max(1, c1["cache_hits"] - c0["cache_hits"] + c1["leaf_evals"] - c0["leaf_evals"])
    else:   # the same games with the evaluation cache: per-network entries, separate launches
        w = _config4_leg(torch, nets, BatchedSelfPlay, blocks, games, sims, ticks, eval_cache=12)
        out["with_eval_cache"] = {k: w[k] for k in ("sims_per_s", "moves_per_s", "ms_per_tick", "leaf_evals_per_tick", "hit_rate")}
    return out


def _eval_cache_leg(torch, nets, BatchedSelfPlay, blocks, games, sims, chunk, headline_sims_per_s, ticks=2400):
    """The headline workload once more with the engine's evaluation cache switched on (spx_config.eval_cache_log2 = 12: 4096 entries
    per game slot): requests for positions a slot has evaluated before are answered from its table inside the fused tick kernel.
    NOT the headline: `value` and `e2e` run one network evaluation per non-terminal simulation, like the reference.  The games are
    identical either way (tests/test_fused_gpu.py); only the number of network passes per move drops."""
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=blocks).eval()
    sp = BatchedSelfPlay(net, game=0, n_games=games, sims=sims, net="tower", seed=0, eval_cache=12)
    e = sp.engine
    e.stagger()
    e.run_ticks(2 * sims, chunk=chunk)
    e.drain_records(); e.drain_results()
    torch.cuda.synchronize()
    c0 = e.counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = 0
    while done < ticks:     # rings drained like the headline loop does between steps
        e.run_ticks(min(sims, ticks - done), chunk=chunk)
        done += min(sims, ticks - done)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    c1 = e.counters()
    d = {k: c1[k] - c0[k] for k in ("sims", "leaf_evals", "cache_hits", "moves")}
    out = {"workload": f"the headline workload ({games} games, {sims} sims/move, ResidualTower-{blocks}) with eval_cache_log2 = 12 (4096 entries of 64 B per game slot); "
                       "NOT the headline: value / e2e evaluate every non-terminal simulation with the network",
           "sims_per_s": d["sims"] / (ms / 1e3), "leaf_evals_per_s": d["leaf_evals"] / (ms / 1e3), "positions_per_s": d["moves"] / (ms / 1e3),
           "hit_rate": d["cache_hits"] / max(1, d["cache_hits"] + d["leaf_evals"]), "ms_per_tick": ms / ticks, "ticks": ticks,
           "speedup_vs_value": d["sims"] / (ms / 1e3) / headline_sims_per_s, "errors": c1["errors"], "records_dropped": c1["records_dropped"]}
    sp.close()
    return out


def _train_leg(torch, nets, blocks, batch=128, steps=100):
    """BASELINE configs[4]'s training half: UpdateWorker.update = 100 SGD steps of batch 128 (updateworker.py:141-149) with the native
    step (csrc/spx_train.cu; synthetic batch resident in HBM), next to PyTorch autograd + torch.optim.SGD on the same module
    (stock settings: TF32 convolutions) for scale."""
    from self_play_reinforcement_learning_b200.train import DeviceTrainer
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=blocks).cuda()
    g = torch.Generator(device="cuda").manual_seed(0)
    boards = torch.randint(-1, 2, (batch, 7, 6), generator=g, device="cuda")
    planes = torch.stack([(boards == 0), (boards == 1), (boards == -1)], 1).float()
    probs = torch.softmax(torch.randn(batch, 7, generator=g, device="cuda"), 1)
    target = torch.rand(batch, generator=g, device="cuda") * 2 - 1
    tr = DeviceTrainer(net, batch_size=batch)

    def timed(fn, n):
        for _ in range(3):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        for _ in range(n):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / n
    l0 = _lib_launches()
    ms = timed(lambda: tr.step(planes, probs, target), steps)
    launches = (_lib_launches() - l0) // (steps + 3)
    loss = tr.step(planes, probs, target).tolist()
    tr.close()
    net.train()
    opt = torch.optim.SGD(net.parameters(), lr=0.01, momentum=0.9, weight_decay=1e-4)

    def torch_step():
        p, v = net.forward_planes(planes)
        l = torch.nn.functional.mse_loss(v.view(-1), target) - (p.log() * probs).sum() / batch
        opt.zero_grad(); l.backward(); opt.step()
    ms_torch = timed(torch_step, 10)
    flop = 3 * batch * flop_per_leaf(blocks)     # forward + backward-data + backward-weights
    return {"workload": f"{steps} SGD steps, batch {batch}, ResidualTower-{blocks}, train mode (batch-statistics BatchNorm, Dropout 0.5), SGD momentum 0.9 wd 1e-4",
            "ms_per_step": ms, "steps_per_s": 1e3 / ms, "ms_per_100_steps": 100 * ms, "kernels_per_step": int(launches), "tflops": flop / (ms / 1e3) / 1e12,
            "loss": loss, "dtype": "tf32 (forward, backward-data) / bf16 (backward-weights) tensor-core GEMMs, fp32 elsewhere",
            "torch_autograd_ms_per_step": ms_torch, "speedup_vs_torch_autograd": ms_torch / ms}


def _config5_leg(torch, nets, blocks, sims, games=1024):
    """BASELINE configs[4] on this GPU: ONE epoch of the reference's loop (self_play_parallel.py:213-291) through
    scheduler.SelfPlayScheduler -- `games` initial games fill the device replay memory, then an epoch of `games` self-play games
    (records stay on the device), 100 native SGD steps of batch 128, the weight refresh and 128 evaluation games against the
    hard-coded OneStepLookahead opponent (main.py:66).  Every game phase is bound by the LATENCY of one game (~36 plies x 800
    ticks), not by throughput; seconds per phase are the scheduler's own (device-synchronised) clock."""
    from self_play_reinforcement_learning_b200.scheduler import SelfPlayScheduler
    torch.manual_seed(0)
    net = nets.ResidualTower(7, 6, 7, num_blocks=blocks).cuda().eval()
    s = SelfPlayScheduler(net, 0, iterations=sims, epoch_length=games, initial_games=games, evaluation_games=128, games_per_gpu=1024,
                          batch_size=128, updates_per_epoch=100, evaluation_opponent="lookahead", save_dir=None)
    import time
    t0 = time.time()
    h = s.train_model(num_epochs=1)[0]
    torch.cuda.synchronize()
    wall = time.time() - t0
    sec = h["seconds"]
    out = {"workload": f"{games} initial + {games} self-play games ({sims} sims/move), 100 SGD steps of batch 128 (native step), weight refresh, "
                       "128 evaluation games vs OneStepLookahead; one GPU; SelfPlayScheduler defaults (evaluation cache on: DESIGN.md 3.9)", "seconds": sec, "wall_s_incl_initial_games": wall,
           "records_in_memory": h["memory"], "loss": h["loss"], "evaluation_reward": h["evaluation_reward"], "trainer": s.trainer_kind,
           "self_play_games_per_s": games / sec["self_play"]}
    # the reference's own epoch size on 8 GPUs is 1500 / 8 = 188 games per GPU: far fewer games than leaf slots, so the phase is
    # bound by the latency of one game.  thread_count = 4 (the reference's default behind its InferenceProxy: virtual loss, 4
    # simulations in flight per tree) makes a move take 200 ticks of 752 leaves instead of 800 ticks of 188.
    from self_play_reinforcement_learning_b200.selfplay import BatchedSelfPlay
    small = {}
    for K, cache in ((1, 0), (4, 0), (1, 12)):   # (1, 12): the sequential search with the evaluation cache: the same games in fewer passes
        sp = BatchedSelfPlay(net, game=0, n_games=188, sims=sims, net="tower", seed=5, games_target=188, search_threads=K, eval_cache=cache)
        torch.cuda.synchronize()
        t0 = time.time()
        while True:
            sp.engine.run_ticks(sp.engine.safe_poll_interval)
            sp.engine.drain_records(); sp.engine.drain_results()
            if sp.engine.all_idle():
                break
        torch.cuda.synchronize()
        small[f"thread_count_{K}" + ("_eval_cache" if cache else "")] = time.time() - t0
        sp.close()
    out["self_play_188_games_seconds"] = small
    out["self_play_188_games_speedup_with_4_threads"] = small["thread_count_1"] / small["thread_count_4"]
    return out


def _lib_launches():
    from self_play_reinforcement_learning_b200 import _lib
    return int(_lib.lib().spx_launch_count())


def _env_roofline(_lib, torch):
    """spx_env_step over 16 Mi Connect4 boards (352 MB of state+inputs, > L2): achieved algorithmic GB/s vs the copy peak.
    43 B/transition = 16 B state read + 16 B written, action 4, player 1, done 1+1, reward 1, valid 2, status 1."""
    import ctypes as C
    n = 1 << 24
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(0)
    state = torch.zeros(n, 2, dtype=torch.int64, device=dev)
    done = torch.zeros(n, dtype=torch.uint8, device=dev)
    reward = torch.zeros(n, dtype=torch.int8, device=dev)
    valid = torch.zeros(n, dtype=torch.int16, device=dev)
    status = torch.zeros(n, dtype=torch.int8, device=dev)
    player = torch.ones(n, dtype=torch.int8, device=dev)
    acts = [torch.randint(0, 7, (n,), generator=g, device=dev, dtype=torch.int32) for _ in range(8)]
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    L = _lib.lib()
    times = []
    for i, a in enumerate(acts):
        torch.cuda.synchronize()
        # events recorded on the launching stream (torch's current stream, passed as `st`) around the single launch
        rec = torch.cuda.Event(enable_timing=True); rec2 = torch.cuda.Event(enable_timing=True)
        rec.record()
        _lib.check(L.spx_env_step(0, n, state.data_ptr(), done.data_ptr(), a.data_ptr(), player.data_ptr(), reward.data_ptr(),
                                  valid.data_ptr(), status.data_ptr(), st), "spx_env_step")
        rec2.record()
        torch.cuda.synchronize()
        if i >= 3:
            times.append(rec.elapsed_time(rec2))
        player = -player
    ms = sum(times) / len(times)
    peaks = _peaks()
    gbs = n * 43 / (ms / 1e3) / 1e9
    return {"kernel": "spx::env_step_quad_kernel<connect4>", "boards": n, "bytes_per_transition": 43, "kernel_ms": ms, "achieved_GBps": gbs,
            "peak_GBps": peaks["hbm_gbs"], "frac": gbs / peaks["hbm_gbs"], "transitions_per_s": n / (ms / 1e3)}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def _ncu_traffic(fused_tick=False):
    """dram bytes of the dominant kernel from the committed `ncu --set full` capture (profiles/): per launch of the tower kernel,
    or per tick of the fused tick kernel."""
    p = os.path.join(ROOT, "profiles", "tower_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))["dram_bytes_per_tick_fused" if fused_tick else "dram_bytes_per_launch"]
        except Exception:
            return None
    return None


if __name__ == "__main__":
    main()

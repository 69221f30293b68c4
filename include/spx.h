/*
 * spx.h -- C ABI of libspx.so: the B200 (sm_100a) batched self-play engine.
 *
 * This is the drop-in boundary for ONE hot path of reubenvanammers/self_play_reinforcement_learning:
 * many concurrent AlphaZero MCTS searches (games/algos/mcts.py) driven by the self-play workers
 * (games/algos/selfplayworker.py, games/algos/self_play_parallel.py).  Every entry point names the
 * reference interface it replaces (file:line relative to the reference tree).
 *
 * Conventions
 *   - plain C types only; every pointer marked "dev" is a CUDA device pointer owned by the caller
 *     (e.g. torch tensor .data_ptr()); "host" pointers are ordinary host memory.
 *   - `stream` is a cudaStream_t passed as void*; calls are asynchronous on it unless stated.
 *   - return value: 0 = OK, <0 = error (SPX_E_*); spx_last_error() gives a thread-local message.
 *   - one engine per device; an engine is not thread-safe (the host serialises calls), exactly like
 *     one reference InferenceWorker/SelfPlayWorker pair.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point returns SPX_E_CUDA.
 *
 * Bitboards: Connect4 bit = col*7 + row (row 0 = bottom, bit 6 of each column never set);
 * TicTacToe bit = x*3 + y (== action index).  "own" = cells equal to +1, "opp" = cells equal to -1
 * in whichever frame the call documents (see oracle/spec.py board_to_bits for the test-side twin).
 */
#ifndef SPX_H
#define SPX_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPX_GAME_CONNECT4 0  /* games/connect4/connect4env.py, default 7x6 only      */
#define SPX_GAME_TICTACTOE 1 /* games/tictactoe/tictactoe_env.py, default 3x3x3 only */

#define SPX_MAX_ACTIONS 9

#define SPX_OPP_MCTS 0
#define SPX_OPP_LOOKAHEAD 1 /* general/hardcoded_players.py:8-37  */
#define SPX_OPP_RANDOM 2    /* general/hardcoded_players.py:40-56 */
#define SPX_OPP_EXTERNAL 3  /* any host-side BasePlayer: moves delivered with spx_set_external_actions */

#define SPX_E_ARG (-1)      /* bad argument                       */
#define SPX_E_CUDA (-2)     /* CUDA runtime error / no device      */
#define SPX_E_STATE (-3)    /* call made in the wrong engine state */
#define SPX_E_OVERFLOW (-4) /* node pool / record ring overflow    */

/* status codes written per game by spx_env_step (the reference raises instead) */
#define SPX_ENV_OK 0
#define SPX_ENV_GAME_OVER (-1)   /* GameOver    connect4env.py:30-31, tictactoe_env.py:24-25 */
#define SPX_ENV_VALUE_ERROR (-2) /* ValueError  connect4env.py:36-37 (full column); also action >= A (IndexError there) */
#define SPX_ENV_SKIPPED (-3)     /* action < 0: slot not stepped                             */

typedef struct spx_engine spx_engine;

/* Mirrors MCTreeSearch.__init__ kwargs (mcts.py:119-136), MCNode constants x=0.25, cpuct=4
 * (mcts.py:24-26) are fixed.  thread_count = search_threads (default 1: the sequential search). */
typedef struct spx_config {
    int32_t game;              /* SPX_GAME_*                                                        */
    int32_t n_games;           /* concurrent game slots on this device (two trees per game)         */
    int32_t sims;              /* `iterations`                                     mcts.py:125,333  */
    int32_t evaluate;          /* MCTreeSearch.evaluate(True): temp/20             mcts.py:273-274  */
    int32_t strong_play;       /* mcts.py:133,307-311                                               */
    int32_t tie_mode;          /* 0: tie noise == 0, 1: counter stream             mcts.py:355      */
    int32_t noise_mode;        /* 0: Dirichlet == 1/A, 1: injected table, 2: generated on device    */
    int32_t emit_records;      /* play_episode(update=...)                 selfplayworker.py:186-190 */
    int32_t max_sims_per_tick; /* terminal re-visits resolved inside one advance call (>=1)         */
    int32_t nodes_per_tree;    /* 0: worst-case bound (sims+1)*ceil(max_moves/2)+max_moves+2        */
    int32_t move_log;          /* 1: keep per-move root statistics for spx_read_move_log (debug)    */
    int32_t two_nets;          /* 1: tree 1 is evaluated by net 1 (compare_models / elo.py:73-91)   */
    int32_t opponent_kind;     /* SPX_OPP_*: 0 MCTS, 1 OneStepLookahead, 2 Random, 3 external (host)  */
    int32_t reserved0;
    double alpha;              /* Dirichlet alpha                                   mcts.py:135      */
    uint64_t seed;             /* counter-stream seed (oracle/spec.py)                              */
    int64_t slot_offset;       /* global index of this device's slot 0 (multi-GPU sharding)         */
    int64_t slot_stride;       /* total slots over all devices; game index = slot + k*stride        */
    int64_t games_target;      /* games with index < target are played, then the slot idles         */
    int64_t record_capacity;   /* Move records buffered on device between drains                    */
    int64_t result_capacity;   /* game results buffered on device between drains                    */
    int32_t search_threads;    /* MCTreeSearch(thread_count=K) behind an InferenceProxy, mcts.py:132,328-331: 0/1 = the
                                * sequential search; K in 2..16 = K search_node tasks in flight per tree with virtual loss and
                                * per-child locks, under the cooperative round-robin schedule.  The leaf batch
                                * then has n_games * K slots (slot of worker k of game g = g * K + k).                  */
    int32_t eval_cache_log2;   /* 0: off.  n in [6, 20]: every game slot keeps a direct-mapped table of 2^n evaluations
                                * {position, network id, weights version} -> (policy, value) and the search answers repeated
                                * requests from it instead of the network (spx_tick_fused; spx_advance after spx_set_eval_cache_versions).  The network is a pure function whose output
                                * does not depend on batch position, so games, records and statistics are unchanged bit for bit;
                                * only spx_counters.leaf_evals drops and cache_hits rises.  64 B per entry.   */
} spx_config;

/* One Move record (mcts.py:17,282-289 + :230): state in the tree's own frame, tree_probs, q, and
 * actual_val stamped at game end (+r for the policy, -r for the opponent). 80 bytes. */
typedef struct spx_record {
    uint64_t own, opp;
    uint64_t game_index;
    float tree_probs[SPX_MAX_ACTIONS];
    float q;
    float actual_val;
    uint8_t tree; /* 0 = policy, 1 = opposing policy (selfplayworker.py:67-90) */
    uint8_t ply;  /* pieces on the board when the move was searched             */
    uint16_t pad0;
    uint64_t pad1;
} spx_record;

/* One result_queue entry {"reward": r, "swap_sides": b} (selfplayworker.py:185). 16 bytes. */
typedef struct spx_result {
    uint64_t game_index;
    int8_t reward;
    uint8_t swap_sides;
    uint8_t plies;
    uint8_t pad[5];
} spx_result;

/* Per-move root statistics (debug / parity): what MCTreeSearch._play sees (mcts.py:272-299). */
typedef struct spx_move_log {
    int32_t tree, ply, action, root_n;
    double root_w;
    int32_t n[SPX_MAX_ACTIONS];
    int32_t pad;
    double w[SPX_MAX_ACTIONS];
    double noise[SPX_MAX_ACTIONS];
} spx_move_log;

typedef struct spx_counters {
    uint64_t sims;          /* completed search_node iterations (mcts.py:333-334)      */
    uint64_t leaf_evals;    /* network evaluations requested (all kinds)               */
    uint64_t terminal_sims; /* sims that ended on a terminal child (no evaluation)     */
    uint64_t path_len_sum;  /* sum over sims of select-path length L                   */
    uint64_t moves;         /* _play calls == Move records produced                    */
    uint64_t games_finished;
    uint64_t nodes_allocated;
    uint64_t ticks;
    uint64_t records_dropped; /* ring overflow (host drained too rarely)               */
    uint64_t errors;          /* node-pool overflow etc.; must stay 0                  */
    uint64_t cache_hits;      /* evaluations answered by the evaluation cache (eval_cache_log2) */
} spx_counters;

const char* spx_last_error(void);
int spx_version(void);
/* number of CUDA kernels launched by this library since load (bench.py "gpu_launches") */
uint64_t spx_launch_count(void);

/* ---------------------------------------------------------------- environment (batched)
 * Replaces Connect4Env.step/get_reward/valid_moves (connect4env.py:29-48,72-92) and
 * TicTacToeEnv.step/get_reward/valid_moves (tictactoe_env.py:23-45,62-82) for n independent boards.
 * state: dev ulonglong2[n] = {own, opp}; done: dev u8[n] in/out (episode_over); action: dev i32[n]
 * (<0 = skip); player: dev i8[n] (+1/-1); reward: dev i8[n] out; valid: dev u16[n] out, bit a set iff
 * action a is legal AFTER the step; status: dev i8[n] out (SPX_ENV_*). */
int spx_env_step(int32_t game, int64_t n, void* state, uint8_t* done, const int32_t* action, const int8_t* player,
                 int8_t* reward, uint16_t* valid, int8_t* status, void* stream);
/* valid_moves() only (connect4env.py:47-48, tictactoe_env.py:42-43) */
int spx_env_valid_moves(int32_t game, int64_t n, const void* state, uint16_t* valid, void* stream);

/* ---------------------------------------------------------------- synthetic network (tests / search-only bench)
 * The oracle/spec.py hash net evaluated on device: policy dev f32[n,A], value dev f32[n]. */
int spx_hashnet_forward(int32_t game, int64_t n, const uint64_t* own, const uint64_t* opp, const uint8_t* needs_eval,
                        const uint8_t* net_id, uint64_t net_seed0, uint64_t net_seed1, float* policy, float* value,
                        void* stream);

/* ---------------------------------------------------------------- search engine
 * Replaces, for n_games concurrent games with two trees each: MCTreeSearch.reset/search/search_node/
 * _expand_node/_play/play_action/_set_node/push_to_queue (mcts.py:166-232,272-367) and
 * SelfPlayer.play_episode (selfplayworker.py:172-224); the InferenceProxy/InferenceWorker queue round
 * trip (inference_proxy.py:21-24, inference_worker.py:89-119) becomes the dense leaf batch below. */
int spx_create(const spx_config* cfg, spx_engine** out);
int spx_destroy(spx_engine* e);
/* (re)start: every slot begins game index slot_offset+slot (generation 0) */
int spx_reset(spx_engine* e, void* stream);
/* injected Dirichlet noise (noise_mode 1): dev f64 [n_table_games][2][table_moves][A], row for game
 * index i is i - first_game_index */
int spx_set_noise_table(spx_engine* e, const double* table, int64_t first_game_index, int64_t n_table_games,
                        int32_t table_moves);
/* One tick: consume the previous tick's network outputs (policy dev f32[n_games,A], value dev
 * f32[n_games]; ignored for slots that did not request an evaluation; may be NULL on the first
 * tick), run every game's state machine until it needs the network again, and fill the leaf batch. */
int spx_advance(spx_engine* e, const float* policy, const float* value, void* stream);
/* engine-owned leaf batch of the last tick, indexed by slot: own/opp in the NET frame
 * (general/modules.py:109-112: the side that just moved is +1), needs_eval u8, net_id u8 */
int spx_leaf_batch(spx_engine* e, uint64_t** own, uint64_t** opp, uint8_t** needs_eval, uint8_t** net_id);
/* root statistics of one tree per slot (root.children[a].n / .w, root.n, root.w): dev outputs */
int spx_root_stats(spx_engine* e, int32_t tree, int32_t* n, double* w, int32_t* root_n, double* root_w,
                   uint16_t* valid, void* stream);
/* copy out and clear buffered records / results (synchronises the stream) */
int spx_drain_records(spx_engine* e, spx_record* host_out, int64_t capacity, int64_t* n_out, void* stream);
int spx_drain_results(spx_engine* e, spx_result* host_out, int64_t capacity, int64_t* n_out, void* stream);
int spx_read_move_log(spx_engine* e, int32_t slot, spx_move_log* host_out, int32_t capacity, int32_t* n_out,
                      void* stream);
int spx_counters_read(spx_engine* e, spx_counters* host_out, void* stream);
/* 1 when every slot is idle (games_target reached) */
int spx_all_idle(spx_engine* e, int32_t* idle_out, void* stream);
int64_t spx_device_bytes(spx_engine* e);
/* (re)start with a new global index for slot 0 and a new games_target (the per-game Policy facade starts every
 * episode this way: one game, then the slot idles) */
int spx_restart(spx_engine* e, int64_t slot_offset, int64_t games_target, void* stream);
/* change the simulations per move (MCTreeSearch.iterations, mcts.py:131) for all following launches; at most the value
 * the engine was created with (the node pool is sized for it).  Searches in progress run on to the new count. */
int spx_set_sims(spx_engine* e, int32_t sims);
/* Test hook: the fused tick kernel's shadow warp does the fp64 PUCT arithmetic (mcts.py:59-84: q = w / n, u = 4 p sqrt(N + 1) /
 * (1 + n), score = player q + u + 1e-6 noise) on the integer pipe, because FP64 instructions next to running tcgen05 MMAs slow
 * the tensor pipe down (csrc/spx_softf64.cuh).  This runs every such operation against the FP64 instruction on n
 * pseudo-random operand sets and returns the number of differing results per operation in mismatches_out[8]
 * (mul, add, div by integer, sqrt of integer, f32 -> f64, uniform from 53 bits, comparisons, scaling / negation): all must be 0. */
int spx_softf64_selftest(uint64_t n, uint64_t seed, uint64_t* mismatches_out, void* stream);
/* Evaluation cache with separate launches (spx_advance + a network forward per tick): tell the engine which weights produce
 * the policy / value arrays the next spx_advance calls consume -- version0 for network 0, version1 for network 1 (two_nets),
 * e.g. spx_tower_version(); 0 = unknown (the default): spx_advance runs without the cache.  Call again after every weight
 * refresh (inference_worker.py:68-73).  spx_tick_fused reads the version from its tower and needs no call. */
int spx_set_eval_cache_versions(spx_engine* e, uint32_t version0, uint32_t version1);
/* SPX_OPP_EXTERNAL: deliver the opposing player's moves, dev i32[n_games] (-1 = none for that slot).  Replaces
 * opposing_policy(s) + policy.play_action(a, -player) in SelfPlayer.get_and_play_moves (selfplayworker.py:206-224). */
int spx_set_external_actions(spx_engine* e, const int32_t* actions, void* stream);
/* dev i32[n_games][6]: {state: 0 running / 1 waiting for an external move / 2 idle, ply, moves played by the policy in
 * this game, the policy's latest action, games finished on this slot, swap_sides} */
int spx_slot_status(spx_engine* e, int32_t* status_out, void* stream);
/* dev i32[n_games]: the tree (0/1) whose evaluation each slot is waiting for, -1 if none (replay logging) */
int spx_pending_tree(spx_engine* e, int32_t* tree_out, void* stream);

/* ---------------------------------------------------------------- network tower (tcgen05)
 * Replaces InferenceWorker.calculate (inference_worker.py:114-119) -> ResidualTower.forward
 * (general/modules.py:88-107, BN in eval mode folded into the convolutions) for the 7x6 Connect4
 * ResidualTower with 128 trunk channels (filter_factor 32) and `num_blocks` residual blocks.
 * Weights arrive as one packed device blob (layout: nets.pack_tower_blob, size spx_tower_blob_bytes), so a
 * weight refresh is a single device-to-device copy (or the receive buffer of an NCCL broadcast). */
typedef struct spx_tower spx_tower;
int64_t spx_tower_blob_bytes(int32_t game, int32_t num_blocks);
int spx_tower_create(int32_t game, int32_t num_blocks, spx_tower** out);
int spx_tower_destroy(spx_tower* t);
/* weight-slice layout the handle expects: 1 = single-CTA kernel, 2 = SM-pair kernel (tcgen05 cta_group::2, default) */
int spx_tower_ncta(spx_tower* t);
/* 1 when the fully connected heads (policy Linear + softmax, value Linear-ReLU-Linear-tanh, modules.py:99-105) run inside the
 * tower kernel (default for the SM-pair kernel; SPX_TOWER_FUSED_HEADS=0 selects the separate heads kernel) */
int spx_tower_fused_heads(spx_tower* t);
/* 1: the conv trunk and the fused value layer compute in fp16 (default: the type the reference's GPU path computes in,
 * torch.cuda.amp.autocast at inference_worker.py:117), 0: bf16 (SPX_TOWER_DTYPE=bf16 at spx_tower_create).  It is the element
 * type spx_tower_load expects the weight stream in. */
int spx_tower_f16(spx_tower* t);
/* a number unique to the weights this tower holds (drawn from a process-wide counter by every spx_tower_load): the tag of the
 * engines' evaluation-cache entries (spx_config.eval_cache_log2) */
uint32_t spx_tower_version(spx_tower* t);
int spx_tower_load(spx_tower* t, const void* dev_blob, int64_t bytes, void* stream);
/* own/opp: dev u64[n] bitboards in the NET frame; needs_eval: dev u8[n] or NULL (all); policy dev f32[n,A]
 * (softmax), value dev f32[n] (tanh).  Rows whose needs_eval is 0 may be left untouched. */
int spx_tower_forward(spx_tower* t, const uint64_t* own, const uint64_t* opp, const uint8_t* needs_eval, int64_t n,
                      float* policy, float* value, void* stream);
/* same, recording caller-owned cudaEvent_t handles (void*) around the conv tower and the heads kernel: bench.py
 * times the dominant kernel on the launching stream with these */
int spx_tower_forward_timed(spx_tower* t, const uint64_t* own, const uint64_t* opp, const uint8_t* needs_eval, int64_t n,
                            float* policy, float* value, void* stream, void* ev_start, void* ev_tower_done, void* ev_end);
/* ---------------------------------------------------------------- TicTacToe network (fp32 CUDA cores)
 * Replaces ConvNetTicTacToe.forward (games/tictactoe/modules.py:55-81, BN in eval mode folded): the "repo's tictactoe net" of
 * BASELINE.json configs[0].  Weights: one fp32 device blob (layout: nets.pack_tttnet_blob, spx_tttnet_blob_floats()). */
typedef struct spx_tttnet spx_tttnet;
int64_t spx_tttnet_blob_floats(void);
int spx_tttnet_create(spx_tttnet** out);
int spx_tttnet_destroy(spx_tttnet* t);
int spx_tttnet_load(spx_tttnet* t, const float* dev_blob, int64_t n_floats, void* stream);
int spx_tttnet_forward(spx_tttnet* t, const uint64_t* own, const uint64_t* opp, const uint8_t* needs_eval, int64_t n, float* policy,
                       float* value, void* stream);

/* ---------------------------------------------------------------- two-network evaluation (compare_models / elo.py:73-91)
 * Route the leaf batch to two networks: rows with net_id == k and needs_eval are packed (stable, by slot) to the front of
 * dense batch k: own2/opp2 dev u64[2][n], needs2 dev u8[2][n], map2 dev i32[2][n] (dense row -> slot).  Each tower then only
 * pays for its own rows (its CTAs skip rows whose needs2 is 0); spx_scatter_outputs copies the results back to the slots. */
int spx_partition_leaves(int64_t n, const uint64_t* own, const uint64_t* opp, const uint8_t* needs_eval, const uint8_t* net_id,
                         uint64_t* own2, uint64_t* opp2, uint8_t* needs2, int32_t* map2, void* stream);
int spx_scatter_outputs(int64_t n, int32_t n_actions, const uint8_t* needs2, const int32_t* map2, const float* policy2, const float* value2,
                        float* policy, float* value, void* stream);

/* ---------------------------------------------------------------- device-resident replay memory
 * Replaces memory_queue -> MCTreeSearch.pull_from_queue (mcts.py:217-222) -> Memory (rl_utils/memory.py:8-30): the Move
 * records never leave HBM.  A bounded FIFO (deque(maxlen=max_size) semantics: appending beyond max_size evicts the oldest),
 * uniform sampling WITHOUT replacement (memory.py:26-30), and the batch assembly of MCTreeSearch.loss (mcts.py:234-243:
 * stacked states -> preprocess planes general/modules.py:115-125, tree_probs, actual_val, q).
 * Logical index 0 is the oldest record.  Host-side bookkeeping (size, head) is exact; none of these calls launches work on
 * an empty memory. */
typedef struct spx_replay spx_replay;
/* max_size: the deque's maxlen; physical_capacity >= max_size: the largest size change_size may ever ask for
 * (UpdateWorker.stagger_memory grows the buffer towards max_mem, updateworker.py:107-109). */
int spx_replay_create(int64_t max_size, int64_t physical_capacity, spx_replay** out);
int spx_replay_destroy(spx_replay* r);
int64_t spx_replay_size(spx_replay* r);     /* len(memory) */
int64_t spx_replay_max_size(spx_replay* r); /* memory.max_size */
int spx_replay_change_size(spx_replay* r, int64_t max_size); /* Memory.change_size (memory.py:22-24): keeps the newest */
int spx_replay_reset(spx_replay* r);                         /* Memory.reset (memory.py:32-33) */
/* Drain the engine's record ring into dev_out (device memory, `capacity` records), sorted by (game_index, tree, ply) so the
 * order does not depend on which warp finished first; *n_out = records written.  Synchronises `stream` (it needs the count). */
int spx_drain_records_device(spx_engine* e, spx_record* dev_out, int64_t capacity, int64_t* n_out, void* stream);
/* Memory.add for n records in device memory, in order (only the last max_size survive) */
int spx_replay_append(spx_replay* r, const spx_record* dev_records, int64_t n, void* stream);
/* copy logical records [first, first+n) to HOST memory (save_memory updateworker.py:123-139, tests) */
int spx_replay_read(spx_replay* r, int64_t first, int64_t n, spx_record* host_out, void* stream);
/* Memory.sample(batch) + the tensor stacking of MCTreeSearch.loss.  Indices: partial Fisher-Yates over the logical index
 * space, draw i from the counter stream (seed, game_uid = step, tree 0, purpose 4, ply 0, sim i, depth 0, idx 0) as
 * j = i + rng_u64 % (size - i)  (integer only: identical on CPU and GPU; oracle/replay.py).  Outputs (device, any may be NULL):
 * idx i64[batch]; boards i64[batch][W][H] (Move.state, +1 own / -1 enemy); planes f32[batch][3][W][H] ([==0, ==+1, ==-1]);
 * tree_probs f32[batch][A]; actual_val f32[batch]; q f32[batch].  batch <= min(size, 4096). */
int spx_replay_sample(spx_replay* r, int32_t game, int64_t batch, uint64_t seed, uint64_t step, int64_t* idx, int64_t* boards,
                      float* planes, float* tree_probs, float* actual_val, float* q, void* stream);

/* Memory.deduplicate("state", ["actual_val", "tree_probs"], Move) (memory.py:47-54, mcts.py:385-386; UpdateWorker option
 * `deduplicate`, updateworker.py:88-89) with the Deduplicator of memory.py:56-94 kept on the device: a persistent table of the
 * distinct states ever folded (first-seen order) with the running f32 sums of tree_probs / actual_val / q and a count; the first
 * call folds the whole current buffer, later calls every record appended since (evicted or not, as the reference's
 * temp_queue); the buffer is then REPLACED by one averaged record per distinct state (sum / count; count in spx_record.pad1,
 * game_index / tree / ply of the first-seen member), the last `maxlen` of them, and `maxlen` becomes the buffer's max_size
 * (maxlen <= 0 = None: the physical capacity).  Sums run in insertion order, so the result equals the reference's bit for bit.
 * As shipped the reference raises TypeError here (Move has a fourth field `q` that create_memory does not fill); q is averaged
 * like the other two value fields.  Synchronises `stream`. */
int spx_replay_deduplicate(spx_replay* r, int64_t maxlen, void* stream);
int64_t spx_replay_unique(spx_replay* r); /* len(deduplicator.counter): distinct states folded so far (0 before the first call) */

/* The whole tick loop in ONE launch (the persistent form of `spx_advance` + `spx_tower_forward` x n_ticks): the tower kernel's
 * CTAs also run the per-game state machine of the games whose leaves they evaluate, tick after tick, without returning to the
 * host -- games are independent, so no grid-wide synchronisation is needed.  Same state, records and outputs as n_ticks
 * separate ticks, bit for bit (tests/test_search_gpu.py).  policy / value: the engine's output buffers (dev f32[n_games][A],
 * f32[n_games]) as passed to spx_advance.  Needs the SM-pair tower with fused heads, one network, same game as the engine. */
int spx_tick_fused(spx_engine* e, spx_tower* t, int32_t n_ticks, float* policy, float* value, void* stream);
/* The same launch, work-conserving: the reference's workers are independent processes that take games off a shared task queue and
 * never wait for each other (self_play_parallel.py:236-253, selfplayworker.py:105-161).  Here the launch holds a budget of
 * n_ticks x ceil(n_games / 14) network passes; every SM pair draws its next tick from that budget when it gets there, so a pair
 * whose games search longer between two evaluations (chains of terminal re-visits) runs fewer ticks instead of making 73 other
 * pairs wait at the end of the launch.  A game's records and results do not depend on how many ticks a launch gives it (every game
 * is replayed bit for bit either way); only the number of ticks per game per launch is no longer exactly n_ticks.  The tick counter
 * advances by n_ticks. */
int spx_tick_fused_balanced(spx_engine* e, spx_tower* t, int32_t n_ticks, float* policy, float* value, void* stream);

/* timing-event helpers (cudaEvent_t behind void*), so hosts without a CUDA binding can time on the launching stream */
int spx_event_create(void** ev_out);
int spx_event_destroy(void* ev);
int spx_event_elapsed_ms(void* ev_start, void* ev_end, float* ms_out); /* synchronises on ev_end */
/* spx_advance with events before/after (HBM roofline of the search kernel) */
int spx_advance_timed(spx_engine* e, const float* policy, const float* value, void* stream, void* ev_start, void* ev_end);

/* ---------------------------------------------------------------- the SGD step on the device
 * Replaces MCTreeSearch.loss + update_from_memory (mcts.py:234-270) as UpdateWorker.update runs them (updateworker.py:141-149,
 * network.train(): BatchNorm on batch statistics, Dropout(0.5) on both head activations, general/modules.py:88-107) with
 * torch.optim.SGD(momentum, weight_decay) (self_play_parallel.py:193) for ResidualTower(7, 6, 7, num_blocks, filter_factor 32)
 * on Connect4 boards.  Convolutions run on tcgen05 in TF32 (fp32 storage and master weights), everything else in fp32.
 * Parameter vector = the tensors of ResidualTower.named_parameters() in order, flattened (general/modules.py:42-75); running
 * statistics = per BatchNorm in module order, running_mean[C] then running_var[C]. */
typedef struct spx_trainer spx_trainer;
int spx_train_create(int32_t num_blocks, int32_t batch, spx_trainer** out);
int spx_train_destroy(spx_trainer* t);
int64_t spx_train_param_count(spx_trainer* t);
int64_t spx_train_running_count(spx_trainer* t);
/* copy parameters / running statistics in (dev pointers, either may be NULL) and rebuild the packed weights;
 * reset_momentum != 0 clears the momentum buffers (a fresh optimizer) */
int spx_train_set_state(spx_trainer* t, const float* dev_params, const float* dev_running, int32_t reset_momentum, void* stream);
/* what: 0 parameters, 1 running statistics, 2 gradients of the last step, 3 momentum buffers -> dev_out */
int spx_train_get_state(spx_trainer* t, int32_t what, float* dev_out, void* stream);
/* One update_from_memory step on a batch in device memory: planes f32[B][3][7][6] (modules.py:115-125, as
 * spx_replay_sample writes them), tree_probs f32[B][7], target f32[B] (actual_val, + q when q_average: mcts.py:243-244),
 * dropout_mask u8[B][2][1344] keep-masks (policy head, value head) or NULL (drawn from (seed, step)); apply_update 0 =
 * forward + backward only.  loss_out (dev, 3 floats or NULL): total, value term, policy term. */
int spx_train_step(spx_trainer* t, const float* planes, const float* tree_probs, const float* target, const uint8_t* dropout_mask,
                   uint64_t seed, uint64_t step, float lr, float momentum, float weight_decay, int32_t apply_update,
                   float* loss_out, void* stream);
/* train-mode network outputs of the last step: probs f32[B][7], value f32[B] (dev, either may be NULL) */
int spx_train_outputs(spx_trainer* t, float* dev_probs, float* dev_value, void* stream);
/* test hook: device pointer / size of an internal plane tensor [C/4][rows][4] (rows = 16 + 56 * batch padded to 16 boards;
 * row = 8 + 56 * board + 7 * col + row_in_col).  which: 0 input, 1 conv output of `layer`, 2 activation of `layer`, 3/4 head conv
 * output / activation, 5/6 trunk gradient ping-pong, 7 conv-output gradient, 8 head conv-output gradient, 9 head activation
 * gradient, 10 identity-branch gradient */
int spx_train_debug_planes(spx_trainer* t, int32_t which, int32_t layer, float** dev_ptr, int64_t* n_floats, int32_t* rows);
/* test hook: dev buffer of 8 int64 receiving clock64() at 8 points of CTA 0 of every conv kernel launch (NULL: off) */
int spx_train_debug_trace(long long* dev_trace);
/* test hook: the backward-weights kernel alone on caller-provided bf16 plane tensors [C/8][rows][8] (x 128 channels, dy N channels,
 * `rows` rows incl. the 16 guard rows); partial_out f32 [S][taps][128 ci][N co]; descriptor strides < 0 = the product's */
int spx_train_debug_wgrad(const void* x, const void* dy, int32_t N, int32_t taps, int32_t rows, int32_t S, float* partial_out,
                          int32_t a_lbo, int32_t a_sbo, int32_t b_lbo, int32_t b_sbo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPX_H */

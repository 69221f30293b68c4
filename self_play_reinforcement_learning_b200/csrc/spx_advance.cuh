// spx_advance.cuh -- the search engine's device side: node-pool layout, per-slot state, and advance_game(), the per-warp state
// machine of one game slot (see the header comment of spx_engine.cu).  Shared by spx_engine.cu (advance_kernel: one warp per
// game) and spx_tower.cu (fused tick kernel: the network CTA that owns a game's leaf also advances the game).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include "spx_common.cuh"
#include "spx_softf64.cuh"

namespace spx {

// ------------------------------------------------------------------------------------------------ layout
template <int GAME> struct NodeLayout {
    static constexpr int A = Rules<GAME>::A;
    static constexpr int OFF_W = 0;
    static constexpr int OFF_N = 8 * A;
    static constexpr int OFF_P = 12 * A;
    static constexpr int OFF_CHILD = 16 * A;
    static constexpr int OFF_META = 20 * A;
    static constexpr int OFF_OWN = 20 * A + 4;
    static constexpr int OFF_OPP = OFF_OWN + 8;
    static constexpr int SIZE = ((OFF_OPP + 8 + 31) / 32) * 32;
    static constexpr int VLS = A <= 8 ? 8 : 16;   // entries per node of the virtual-loss side array (threaded search)
    static_assert(OFF_OWN % 8 == 0, "bitboards must be 8-byte aligned");
};

enum { CHILD_UNEXPANDED = -1, CHILD_TERM_DRAW = -2, CHILD_TERM_WIN = -3 };
enum { PH_IDLE = 0, PH_RESET = 1, PH_SEARCH = 2, PH_REROOT = 3, PH_ENVSTEP = 4 };
enum { PK_NONE = 0, PK_ROOT = 1, PK_EXPAND = 2, PK_REROOT = 3, PK_THREADS = 4 };   // PK_THREADS: the pending evaluations are the workers' (threaded search)
#define SPX_MAX_PATH 64
#define SPX_MAX_OWN_MOVES 22
#define SPX_MAX_PLIES 44

struct TreeState {   // 32 bytes = two 16-byte words (ld_tree / st_tree)
    double root_w;
    int root, root_n, root_player, moves_played, node_count, n_rec;
};

// One game slot.  Laid out in 16-byte groups so that advance_prefetch() moves it with a dozen 128-bit loads; advance_game()
// keeps the fields it needs in registers (no dynamically indexed copy: the first version held `GameState s = *gp` with
// `s.tree[T]`, a 496-byte local-memory frame on the latency-critical chain) and writes them back with 128-bit stores.
struct alignas(16) GameState {
    int phase, sub_tree, mover_tree, sims_done;                                  //   0
    int ply, swap, last_action, n_moves_logged;                                  //  16
    int pend_kind, pend_tree, pend_parent, pend_action;                          //  32  the evaluation this slot waits for
    int pend_depth, pend_parent_player, leaf_waiting, pad1;                      //  48  leaf_waiting: emitted, not yet evaluated
    u64 pend_own, pend_opp;                                                      //  64  its child state, TREE frame
    u64 env_own, env_opp;                                                        //  80  env frame: own = +1 = the policy (tree 0)
    u64 game_index; int root_vl[2];                                              //  96  root_vl: MCNode.virtual_loss of the two roots (threaded search)
    TreeState tree[2];                                                           // 112
    u64 cnt_sims, cnt_evals, cnt_term, cnt_path, cnt_moves, cnt_games, cnt_nodes, cnt_err;   // 176 (updated with RED.ADD)
    u64 cnt_hits, pad2;                                                          // 240  evaluations answered by the slot's evaluation cache
};
static_assert(sizeof(TreeState) == 32 && sizeof(GameState) == 256, "GameState layout");

// One worker thread of the threaded search (MCTreeSearch(thread_count = K), mcts.py:328-331): idle, or blocked in the network
// call of the child (parent, action) it expands -- whose lock it holds -- with its select path in wpaths[g][k][].
struct alignas(16) Worker {
    int kind, parent, action, depth;      // kind: PK_NONE idle / PK_EXPAND waiting for its evaluation
    int pplayer, pad0, pad1, pad2;
    u64 own, opp;                         // the child's state, tree frame
};
static_assert(sizeof(Worker) == 48, "Worker layout");
#define SPX_LOCK_SHIFT 17                 // node meta: bits 0..8 valid moves, bit 16 player, bits 17..25 "child is locked" (mcts.py:47,86-88)

struct EngineDev {
    spx_config cfg;
    int nodes_per_tree;
    int K;                 // leaf slots per game: 1, or cfg.search_threads (slot of worker k of game g = g * K + k)
    Worker* workers;       // [G][K]                       (K > 1)
    unsigned* wpaths;      // [G][K][SPX_MAX_PATH]         (K > 1)
    unsigned short* vlpool;  // [G][2][nodes_per_tree][8 or 16]: MCNode.virtual_loss of a node's children (K > 1)
    GameState* games;
    char* pool;
    unsigned* paths;       // [G][SPX_MAX_PATH]  node<<4 | action
    double* noise;         // [G][SPX_MAX_ACTIONS] Dirichlet noise of the search in progress
    spx_record* temp_rec;  // [G][2][SPX_MAX_OWN_MOVES]
    spx_move_log* mlog;    // [G][SPX_MAX_PLIES] or null
    spx_record* rec_ring; unsigned long long* rec_count; unsigned long long* rec_dropped;
    spx_result* res_ring; unsigned long long* res_count;
    const double* noise_table; long long table_first, table_games; int table_moves;
    u64* leaf_own; u64* leaf_opp; unsigned char* needs_eval; unsigned char* net_id;
    unsigned long long* ticks;
    int* ext_action;       // [G] external opponent's next move (>= 0) or -1 (opponent_kind == SPX_OPP_EXTERNAL)
    int* own_action;       // [G][2]: {number of moves the policy (tree 0) has played this game, its latest action}
    // evaluation cache (cfg.eval_cache_log2 > 0; set per launch by spx_tick_fused, null everywhere else): per slot a direct-mapped
    // table of 64-byte entries {own, opp (net frame) | tag, value, policy[0..1] | policy[2..5] | policy[6..8], 0}
    uint4* ecache;         // [G][1 << ecache_log2][4]
    unsigned ecache_log2;
    const double* sqrt_table;   // [sf::SQRT_TABLE] correctly rounded sqrt(m), for the integer-pipe arithmetic of the fused kernel's shadow warp
    unsigned ecache_tag[2];   // per network id: weights version << 8 | id << 1 | 1: entries of other weights never match
};

}  // namespace spx

struct spx_engine {
    spx::EngineDev d;
    int device;
    int64_t bytes;
    uint4* ecache;         // the evaluation cache's tables (cfg.eval_cache_log2 > 0), handed to the kernels per launch
    unsigned cache_ver[2]; // spx_set_eval_cache_versions: the weights versions behind spx_advance's policy / value inputs (0 = unknown: no cache)
};

namespace spx {

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// n**k, correctly rounded (twin of ox_pow_int_exact): np.power(n, 1/temp) for integral 1/temp (mcts.py:100-101).
// 640-bit integer in ten 64-bit limbs that stay in registers (every limb index is a compile-time constant: no local-memory
// array); n**20 fits for every 32-bit n.  Round to nearest even on the top 53 bits.
__host__ __device__ inline double pow_int_exact(unsigned n, int k) {
    if (k == 0) return 1.0;
    if (n == 0) return 0.0;
    constexpr int NL = 10;
    u64 l0 = 1, l1 = 0, l2 = 0, l3 = 0, l4 = 0, l5 = 0, l6 = 0, l7 = 0, l8 = 0, l9 = 0;
    for (int i = 0; i < k; ++i) {
        u64 carry = 0;
#ifdef __CUDA_ARCH__
#define SPX_MULADD(L) { const u64 lo = L * (u64)n, hi = __umul64hi(L, (u64)n); L = lo + carry; carry = hi + (L < lo ? 1ULL : 0ULL); }
#else
#define SPX_MULADD(L) { const unsigned __int128 t = (unsigned __int128)L * n + carry; L = (u64)t; carry = (u64)(t >> 64); }
#endif
        SPX_MULADD(l0) SPX_MULADD(l1) SPX_MULADD(l2) SPX_MULADD(l3) SPX_MULADD(l4)
        SPX_MULADD(l5) SPX_MULADD(l6) SPX_MULADD(l7) SPX_MULADD(l8) SPX_MULADD(l9)
#undef SPX_MULADD
        if (carry) return INFINITY;
    }
    // the two top non-zero limbs (hi, lo), whether anything below them is set, and the index of the top limb
    u64 hi = l0, lo = 0, rest = 0;
    int top = 0;
#define SPX_TOP(J, L, LM1, BELOW) if (L) { hi = L; lo = LM1; rest = BELOW; top = J; }
    SPX_TOP(1, l1, l0, 0ULL)
    SPX_TOP(2, l2, l1, l0)
    SPX_TOP(3, l3, l2, l0 | l1)
    SPX_TOP(4, l4, l3, l0 | l1 | l2)
    SPX_TOP(5, l5, l4, l0 | l1 | l2 | l3)
    SPX_TOP(6, l6, l5, l0 | l1 | l2 | l3 | l4)
    SPX_TOP(7, l7, l6, l0 | l1 | l2 | l3 | l4 | l5)
    SPX_TOP(8, l8, l7, l0 | l1 | l2 | l3 | l4 | l5 | l6)
    SPX_TOP(9, l9, l8, l0 | l1 | l2 | l3 | l4 | l5 | l6 | l7)
#undef SPX_TOP
    (void)NL;
#ifdef __CUDA_ARCH__
    const int lz = __clzll((long long)hi);
#else
    const int lz = __builtin_clzll(hi);
#endif
    const int nbits = top * 64 + 64 - lz;
    if (nbits <= 53) return (double)hi;                       // top == 0: exact
    // normalise the 128-bit window so that its top bit is bit 127
    const u64 nh = lz ? ((hi << lz) | (lo >> (64 - lz))) : hi, nl = lz ? (lo << lz) : lo;
    u64 mant = nh >> 11;                                      // 53 bits
    const int half = (int)((nh >> 10) & 1ULL);
    const int sticky = ((nh & 0x3FFULL) | nl | rest) != 0;
    if (half && (sticky || (mant & 1ULL))) mant += 1;         // may carry to 2^53: the scaling below is exact either way
    return ldexp((double)mant, nbits - 53);
}

// Gamma(alpha,1) variate from the counter stream (same draw schedule as oracle gamma_variate; CUDA libm,
// so only statistically -- not bitwise -- equal to the CPU twin: parity tests inject noise tables).
__device__ inline double gamma_variate(u64 prefix, int action, double alpha) {
    unsigned attempt = 0;
#define SPX_U(j) rng_uniform_from(prefix, attempt, 0, (u64)action * 4 + (j))
    if (alpha == 1.0) return -log(1.0 - SPX_U(0));
    if (alpha < 1.0) {
        for (;; ++attempt) {
            double U = SPX_U(0), V = -log(1.0 - SPX_U(1));
            if (U <= 1.0 - alpha) { double X = pow(U, 1.0 / alpha); if (X <= V) return X; }
            else { double Y = -log((1.0 - U) / alpha); double X = pow(1.0 - alpha + alpha * Y, 1.0 / alpha); if (X <= V + Y) return X; }
            if (attempt > 1000) return 1e-300;
        }
    }
    double b = alpha - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * b);
    for (;; ++attempt) {
        double X = sqrt(-2.0 * log(1.0 - SPX_U(0))) * cospi(2.0 * SPX_U(1));   // cospi: no large-argument reduction path (a local-memory table)
        double V = 1.0 + cc * X;
        if (attempt > 1000) return b;
        if (V <= 0.0) continue;
        V = V * V * V;
        double U = SPX_U(2);
        if (U < 1.0 - 0.0331 * (X * X) * (X * X)) return b * V;
        if (log(U) < 0.5 * X * X + b * (1.0 - V + log(V))) return b * V;
    }
#undef SPX_U
}

#ifndef SPX_PREFETCH
#define SPX_PREFETCH "prefetch.global.L2"   // (.L1 measured: no gain -- 0.4661 vs 0.4628 ms per fused tick, within box noise)
#endif
template <int GAME> struct Ctx {
    typedef Rules<GAME> R;
    typedef NodeLayout<GAME> L;
    const EngineDev& E;
    int g, lane;
    char* tbase;  // node pool of the tree currently worked on
    unsigned short* vbase;   // its virtual-loss side array (threaded search only)
    __device__ Ctx(const EngineDev& e, int g_, int lane_) : E(e), g(g_), lane(lane_), tbase(nullptr), vbase(nullptr) {}
    __device__ __forceinline__ void use_tree(int t) {
        tbase = E.pool + ((size_t)g * 2 + t) * (size_t)E.nodes_per_tree * L::SIZE;
        if (E.vlpool) vbase = E.vlpool + ((size_t)g * 2 + t) * (size_t)E.nodes_per_tree * L::VLS;
    }
    __device__ __forceinline__ char* node(int idx) const { return tbase + (size_t)idx * L::SIZE; }
};

// terminal value v = r (already multiplied by the mover) or the strong_play formula (mcts.py:305-313)
__device__ __forceinline__ double terminal_value(int strong, int r_signed, u64 parent_own, u64 parent_opp) {
    if (!strong) return r_signed > 0 ? 1.0 : (r_signed < 0 ? -1.0 : 0.0);   // (double)r_signed without a conversion instruction
    int num_steps = __popcll(parent_own | parent_opp) + 1;
    return __dmul_rn(__dsub_rn(1.18, __ddiv_rn((double)(9 * num_steps), 350.0)), (double)r_signed);
}

// w += v; n += 1 on every edge of the path and on the root itself (MCNode.backup, mcts.py:94-98;
// ancestors above the current root are never read again so they are skipped).
template <int GAME, bool SOFT = false>
__device__ __forceinline__ void backup_path(const Ctx<GAME>& c, unsigned p0, unsigned p1, int depth, double v, TreeState& ts) {
    typedef NodeLayout<GAME> L;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        int d = c.lane + 32 * half;
        if (d < depth) {
            unsigned e = half ? p1 : p0;
            char* nd = c.node((int)(e >> 4));
            int a = (int)(e & 15u);
            int* pn = (int*)(nd + L::OFF_N) + a;
            double* pw = (double*)(nd + L::OFF_W) + a;
            *pn = *pn + 1;
            *pw = FP<SOFT>::add(*pw, v);
        }
    }
    ts.root_n += 1;
    ts.root_w = FP<SOFT>::add(ts.root_w, v);
}

// create_children (mcts.py:103-107) for a freshly evaluated position + link from its parent edge
template <int GAME>
__device__ __forceinline__ int alloc_node(const Ctx<GAME>& c, TreeState& ts, u64 own, u64 opp, int player, float my_p) {
    typedef NodeLayout<GAME> L;
    typedef Rules<GAME> R;
    int idx = ts.node_count;
    if (idx >= c.E.nodes_per_tree) return -1;
    ts.node_count = idx + 1;
    char* nd = c.node(idx);
    if (c.lane < R::A) {
        ((double*)(nd + L::OFF_W))[c.lane] = 0.0;
        ((int*)(nd + L::OFF_N))[c.lane] = 0;
        ((float*)(nd + L::OFF_P))[c.lane] = my_p;
        ((int*)(nd + L::OFF_CHILD))[c.lane] = CHILD_UNEXPANDED;
    }
    if (c.lane == 0) {
        *(unsigned*)(nd + L::OFF_META) = valid_mask<GAME>(own, opp) | (player < 0 ? 0x10000u : 0u);
        *(u64*)(nd + L::OFF_OWN) = own;
        *(u64*)(nd + L::OFF_OPP) = opp;
    }
    if (c.vbase && c.lane < L::VLS) c.vbase[(size_t)idx * L::VLS + c.lane] = 0;   // fresh MCNodes: virtual_loss = 0 (mcts.py:43)
    return idx;
}

// ------------------------------------------------------------------------------------------------ slot state in registers
__device__ __forceinline__ TreeState ld_tree(const TreeState* t) {
    const int4 x = reinterpret_cast<const int4*>(t)[0], y = reinterpret_cast<const int4*>(t)[1];
    TreeState r;
    r.root_w = __hiloint2double(x.y, x.x); r.root = x.z; r.root_n = x.w;
    r.root_player = y.x; r.moves_played = y.y; r.node_count = y.z; r.n_rec = y.w;
    return r;
}
__device__ __forceinline__ void st_tree(TreeState* t, const TreeState& r) {
    reinterpret_cast<int4*>(t)[0] = make_int4(__double2loint(r.root_w), __double2hiint(r.root_w), r.root, r.root_n);
    reinterpret_cast<int4*>(t)[1] = make_int4(r.root_player, r.moves_played, r.node_count, r.n_rec);
}

// The slot's state as advance_game wants it, loaded with 128-bit loads that every lane issues for the same addresses (one
// transaction each).  Callers that have something to wait for (the fused tick kernel: the network outputs of this very leaf)
// issue advance_prefetch() BEFORE the wait, so the state is in registers when the outputs arrive.
struct AdvPre {
    int4 a, b, c, d;
    ulonglong2 pend;
    u64 game_index;
    TreeState t0, t1;
    unsigned p0, p1;   // this lane's entries of the pending path (lane d / d + 32)
    double noise;      // this lane's Dirichlet noise of the search in progress
};
template <int GAME>
__device__ __forceinline__ AdvPre advance_prefetch(const EngineDev& E, const int g, const int lane) {
    const GameState* gp = E.games + g;
    const int4* q = reinterpret_cast<const int4*>(gp);
    AdvPre r;
    r.a = q[0]; r.b = q[1]; r.c = q[2]; r.d = q[3];
    r.pend = reinterpret_cast<const ulonglong2*>(gp)[4];
    r.game_index = gp->game_index;
    r.t0 = ld_tree(&gp->tree[0]); r.t1 = ld_tree(&gp->tree[1]);
    const unsigned* pp = E.paths + (size_t)g * SPX_MAX_PATH;
    r.p0 = pp[lane]; r.p1 = pp[lane + 32];
    r.noise = lane < Rules<GAME>::A ? E.noise[(size_t)g * SPX_MAX_ACTIONS + lane] : 0.0;
    return r;
}

// ------------------------------------------------------------------------------------------------ threaded search, one tick
// MCTreeSearch.search with thread_count = K behind an InferenceProxy (mcts.py:328-331) under the cooperative round-robin schedule
// (tests/golden/threaded.json holds the unmodified reference forced into it): worker k
// resumes with the evaluation it was blocked on (create_children, backup, virtual loss removed, lock released), then runs
// search_node tasks until it blocks in the next network call; tasks that need none (terminal leaf, "all states in use") complete
// on the spot.  select_prob sees the other workers' virtual losses ((w - vl) / (n + vl), sqrt(parent.n + parent.vl), 1 + n + vl:
// mcts.py:59-78) and skips locked children (:86-88).  Leaf of worker k -> leaf slot g * K + k.  Returns true when every task
// has run and no worker is waiting.  One warp; `ts` / `root_vl` are the searching tree's state (registers of the caller).
template <int GAME>
__device__ __noinline__ bool search_round_threaded(const EngineDev& E, Ctx<GAME>& c, GameState* gp, const int g, const int lane, const int T,
                                                   TreeState& ts, int& root_vl, int& sims_started, const u64 game_index, const int ply,
                                                   const double my_noise, const float* __restrict__ policy_in, const float* __restrict__ value_in) {
    typedef Rules<GAME> R;
    typedef NodeLayout<GAME> L;
    typedef unsigned long long ull;
    constexpr int A = R::A;
    const spx_config& cfg = E.cfg;
    const int K = E.K, ls = g * K;
    const u64 tie_pre = rng_prefix(cfg.seed, game_index, T, PURPOSE_TIE, ply);
    unsigned short* vlp = c.vbase;
    auto count = [&](u64* counter, const int by) { if (lane == 0 && by) atomicAdd((ull*)counter, (ull)by); };
    // virtual loss of the nodes of a finished task's node_list: the root and the inner nodes (children reached by edges 0 .. depth-2)
    auto remove_vl = [&](const unsigned p0, const unsigned p1, const int depth) {
        root_vl -= 1;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int d = lane + 32 * half;
            if (d < depth - 1) {
                const unsigned e = half ? p1 : p0;
                unsigned short* pv = vlp + (size_t)(e >> 4) * L::VLS + (e & 15u);
                *pv = (unsigned short)(*pv - 1);
            }
        }
    };
    int waiting = 0;
    for (int k = 0; k < K; ++k) {
        Worker* wk = E.workers + (size_t)ls + k;
        unsigned* wp = E.wpaths + ((size_t)ls + k) * SPX_MAX_PATH;
        if (wk->kind == PK_EXPAND) {
            // ---- resume: _expand_node after the network call (mcts.py:316-320), backup, remove_virtual_loss, lock.release (:360-364)
            const int parent = wk->parent, action = wk->action, depth = wk->depth, pplayer = wk->pplayer;
            const u64 own = wk->own, opp = wk->opp;
            const float my_p = lane < A ? policy_in[((size_t)ls + k) * A + lane] : 0.f;
            const double v = __dmul_rn((double)value_in[ls + k], (double)pplayer);
            const int idx = alloc_node<GAME>(c, ts, own, opp, -pplayer, my_p);
            if (idx < 0) count(&gp->cnt_err, 1);
            count(&gp->cnt_nodes, 1);
            if (lane == 0) {
                char* pn = c.node(parent);
                if (idx >= 0) ((int*)(pn + L::OFF_CHILD))[action] = idx;
                *(unsigned*)(pn + L::OFF_META) &= ~(1u << (SPX_LOCK_SHIFT + action));
                wk->kind = PK_NONE;
            }
            const unsigned p0 = lane < depth ? wp[lane] : 0u, p1 = lane + 32 < depth ? wp[lane + 32] : 0u;
            __syncwarp();
            backup_path<GAME>(c, p0, p1, depth, v, ts);
            remove_vl(p0, p1, depth);
            __syncwarp();
        }
        bool blocked = false;
        while (!blocked && sims_started < cfg.sims) {
            const int sim = sims_started++;
            // ---------------- search_node (mcts.py:340-367) with virtual loss
            int node = ts.root, N = ts.root_n, player = ts.root_player, depth = 0;
            root_vl += 1;
            int VL = root_vl;
            unsigned p0 = 0, p1 = 0;
            int child = 0, act = 0;
            bool all_bad = false;
            unsigned meta_u = 0;
            for (;;) {
                const char* nd = c.node(node);
                double score = -INFINITY, w = 0.0;
                int ch = 0, n = 0, vl = 0;
                float p = 0.f;
                unsigned meta = 0;
                if (lane < A) {
                    w = ((const double*)(nd + L::OFF_W))[lane];
                    n = ((const int*)(nd + L::OFF_N))[lane];
                    p = ((const float*)(nd + L::OFF_P))[lane];
                    ch = ((const int*)(nd + L::OFF_CHILD))[lane];
                    meta = *(const unsigned*)(nd + L::OFF_META);
                    vl = (int)vlp[(size_t)node * L::VLS + lane];
                }
                const double sqrt_n = __dsqrt_rn((double)(N + VL));
                const double tie = cfg.tie_mode ? __dmul_rn(0.000001, rng_uniform_from(tie_pre, (unsigned)sim, (unsigned)depth, (u64)lane)) : 0.0;
                bool ok = false;
                if (lane < A) {
                    ok = ((meta >> lane) & 1u) && !((meta >> (SPX_LOCK_SHIFT + lane)) & 1u);                 // valid: :86-88
                    if (ok) {
                        const int n_eff = n + vl;
                        const double q = n_eff ? __ddiv_rn(__dsub_rn(w, (double)vl), (double)n_eff) : 0.0;      // :59-62
                        double p_eff = (double)p;
                        if (depth == 0) p_eff = __dadd_rn(__dmul_rn(my_noise, 0.25), __dmul_rn((double)p, 0.75));
                        const double u = __ddiv_rn(__dmul_rn(__dmul_rn(4.0, p_eff), sqrt_n), (double)(1 + n + vl));   // :71-78
                        score = __dadd_rn(__dmul_rn((double)player, q), u);
                    } else score = -10000000000.0;
                }
                if (!__ballot_sync(0xffffffffu, lane < A && !(score < -100000.0))) { all_bad = true; break; }   // :349-354 "all states in use"
                if (lane < A && cfg.tie_mode) score = __dadd_rn(score, tie);
                int best = lane;
                double bs = score;
#pragma unroll
                for (int off = (A <= 8 ? 4 : 8); off > 0; off >>= 1) {
                    const double os = __shfl_xor_sync(0xffffffffu, bs, off);
                    const int ob = __shfl_xor_sync(0xffffffffu, best, off);
                    if (os > bs || (os == bs && ob < best)) { bs = os; best = ob; }
                }
                best = __shfl_sync(0xffffffffu, best, 0);
                child = __shfl_sync(0xffffffffu, ch, best);
                const int n_edge = __shfl_sync(0xffffffffu, n, best), vl_edge = __shfl_sync(0xffffffffu, vl, best);
                meta_u = __shfl_sync(0xffffffffu, meta, 0);
                const unsigned entry = ((unsigned)node << 4) | (unsigned)best;
                if (lane == (depth & 31)) { if (depth < 32) p0 = entry; else p1 = entry; }
                depth += 1;
                act = best;
                if (child < 0 || depth >= SPX_MAX_PATH) break;
                // descend: the child joins node_list, its virtual loss goes up before ITS children are scored (:344-345)
                if (lane == 0) vlp[(size_t)node * L::VLS + best] = (unsigned short)(vl_edge + 1);
                __syncwarp();
                node = child; N = n_edge; VL = vl_edge + 1; player = -player;
            }
            if (all_bad) { count(&gp->cnt_sims, 1); continue; }   // the task is over; its virtual losses stay (as in the reference)
            count(&gp->cnt_path, depth);
            // ---------------- the leaf (node, act): lock (:358), _expand_node (:301-321)
            char* pnd = c.node(node);
            if (lane == 0) *(unsigned*)(pnd + L::OFF_META) = meta_u | (1u << (SPX_LOCK_SHIFT + act));
            const u64 par_own = *(const u64*)(pnd + L::OFF_OWN), par_opp = *(const u64*)(pnd + L::OFF_OPP);
            u64 c_own = par_own, c_opp = par_opp;
            int r = 0, done = 0;
            if (child == CHILD_UNEXPANDED) env_step<GAME>(c_own, c_opp, act, player, r, done);
            else { done = 1; r = (child == CHILD_TERM_WIN); }
            count(&gp->cnt_sims, 1);
            if (done) {
                const double v = terminal_value(cfg.strong_play, r * player, par_own, par_opp);
                __syncwarp();
                if (lane == 0) {
                    if (child == CHILD_UNEXPANDED) ((int*)(pnd + L::OFF_CHILD))[act] = r ? CHILD_TERM_WIN : CHILD_TERM_DRAW;
                    *(unsigned*)(pnd + L::OFF_META) = meta_u;      // lock.release
                }
                __syncwarp();
                backup_path<GAME>(c, p0, p1, depth, v, ts);
                remove_vl(p0, p1, depth);
                __syncwarp();
                count(&gp->cnt_term, 1);
                continue;
            }
            // blocks in self.network(s, parent.player): the leaf goes to slot ls + k, the worker keeps lock and virtual losses
            if (lane < depth) wp[lane] = p0;
            if (lane + 32 < depth) wp[lane + 32] = p1;
            if (lane == 0) {
                wk->kind = PK_EXPAND; wk->parent = node; wk->action = act; wk->depth = depth; wk->pplayer = player;
                wk->own = c_own; wk->opp = c_opp;
                E.leaf_own[ls + k] = player > 0 ? c_own : c_opp;
                E.leaf_opp[ls + k] = player > 0 ? c_opp : c_own;
                E.needs_eval[ls + k] = 1;
                E.net_id[ls + k] = (unsigned char)(cfg.two_nets ? T : 0);
                atomicAdd((ull*)&gp->cnt_evals, 1ULL);
            }
            __syncwarp();
            blocked = true;
        }
        if (blocked) waiting += 1;
        else if (lane == 0) E.needs_eval[ls + k] = 0;
    }
    return waiting == 0 && sims_started >= cfg.sims;
}

// ------------------------------------------------------------------------------------------------ evaluation cache
// The network is a pure function of (weights, position) and its output is bitwise independent of batch position, so an
// evaluation of a position the slot has evaluated before (the two trees of a game search overlapping subtrees, Connect4 move
// orders transpose, every game on the slot starts from the same opening: 46-52 % of all requests at 800 sims/move,
// scripts/dbg_transpositions.py) can be answered from a table without changing a single bit of the game.  One warp owns a slot at
// any time, so the table needs no atomics; it is read and written through L2 (.cg) only.
__device__ __forceinline__ unsigned ecache_slot(const EngineDev& E, u64 own, u64 opp, int net) {
    u64 h = (own * 0x9E3779B97F4A7C15ULL) ^ (opp * 0xC2B2AE3D27D4EB4FULL) ^ (u64)(net * 0x632BE5AB);
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ULL; h ^= h >> 32;
    return (unsigned)h & ((1u << E.ecache_log2) - 1u);
}
__device__ __forceinline__ uint4* ecache_entry(const EngineDev& E, int g, unsigned slot) {
    return E.ecache + ((((size_t)g << E.ecache_log2) + slot) << 2);
}
// all 32 lanes; true = hit: p_out = this lane's prior (lane < A), v_out = the value
template <int A>
__device__ __forceinline__ bool ecache_lookup(const EngineDev& E, int g, int lane, u64 own, u64 opp, int net, float& p_out, float& v_out) {
    const uint4* ent = ecache_entry(E, g, ecache_slot(E, own, opp, net));
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (lane < 4) w = __ldcg(ent + lane);
    const unsigned tag = E.ecache_tag[net & 1];
    bool ok = true;
    if (lane == 0) ok = w.x == (unsigned)own && w.y == (unsigned)(own >> 32) && w.z == (unsigned)opp && w.w == (unsigned)(opp >> 32);
    if (lane == 1) ok = w.x == tag;
    if ((__ballot_sync(0xffffffffu, ok) & 3u) != 3u) return false;
    v_out = __uint_as_float(__shfl_sync(0xffffffffu, w.y, 1));
    const int f = 6 + (lane < A ? lane : 0), src = f >> 2, comp = f & 3;   // float f of the entry = policy[lane]
    const unsigned x = __shfl_sync(0xffffffffu, w.x, src), y = __shfl_sync(0xffffffffu, w.y, src);
    const unsigned z = __shfl_sync(0xffffffffu, w.z, src), q = __shfl_sync(0xffffffffu, w.w, src);
    p_out = lane < A ? __uint_as_float(comp == 0 ? x : comp == 1 ? y : comp == 2 ? z : q) : 0.f;
    return true;
}
// all 32 lanes: my_p = this lane's prior as the network returned it
template <int A>
__device__ __forceinline__ void ecache_insert(const EngineDev& E, int g, int lane, u64 own, u64 opp, int net, float my_p, float v) {
    uint4* ent = ecache_entry(E, g, ecache_slot(E, own, opp, net));
    unsigned vals[4];
#pragma unroll
    for (int cidx = 0; cidx < 4; ++cidx) {
        const int f = 4 * (lane & 3) + cidx;
        const unsigned pv = __shfl_sync(0xffffffffu, __float_as_uint(my_p), (f - 6) & 31);
        vals[cidx] = f == 0 ? (unsigned)own : f == 1 ? (unsigned)(own >> 32) : f == 2 ? (unsigned)opp : f == 3 ? (unsigned)(opp >> 32)
                   : f == 4 ? E.ecache_tag[net & 1] : f == 5 ? __float_as_uint(v) : (f - 6 < A ? pv : 0u);
    }
    if (lane < 4) __stcg(ent + lane, make_uint4(vals[0], vals[1], vals[2], vals[3]));
}

enum { ADV_EMITTED = 1, ADV_IDLE = 2, ADV_PARKED = 4 };   // advance_game's result; 0 = the sim budget ran out before a leaf came up

// ------------------------------------------------------------------------------------------------ one game, one tick
// One game slot's share of a tick, executed by one whole warp (all 32 lanes call it together): consume the evaluation the slot
// asked for (my_p = this lane's prior, v_in = the value; read by the caller), run its state machine until the next network
// request or until `budget` simulations ended without one (terminal re-visits need no network), publish the leaf.
// Called by advance_kernel (one warp per game) and by the fused tick kernel of spx_tower.cu.
//   defer_leaf: the leaf will not be evaluated by the tick that follows this call (the fused kernel's shadow warp works ahead):
//   needs_eval stays 0 and the slot is marked leaf_waiting; the next advance_game on the slot hands the leaf out instead of
//   advancing.   count_tick: bump the engine's tick counter (slot 0 of advance_kernel).
// Registers: only what the select loop needs stays live through the function (phase words, game index, ONE tree's TreeState,
// this lane's noise); the pending-evaluation words, the env boards, the other tree and the counters live in the slot's
// GameState and are read / written (lane 0; counters with RED.ADD) where the state machine touches them.
// SOFT: the PUCT / backup arithmetic runs on the integer pipe (spx_softf64.cuh: same bits, no FP64 instruction next to the MMAs)
template <int GAME, bool THREADED = false, bool CACHE = false, bool SOFT = false>
__device__ __forceinline__ int advance_game(const EngineDev& E, const int g, const int lane, const AdvPre& pre, const float my_p,
                                            const float v_in, int budget, const bool defer_leaf, const bool count_tick,
                                            u64& out_own, u64& out_opp, const float* __restrict__ policy_in = nullptr,
                                            const float* __restrict__ value_in = nullptr) {
    typedef Rules<GAME> R;
    typedef NodeLayout<GAME> L;
    typedef unsigned long long ull;
    typedef FP<SOFT> F;
    constexpr int A = R::A;
    Ctx<GAME> c(E, g, lane);
    GameState* gp = E.games + g;
    const spx_config& cfg = E.cfg;
    const int ls = THREADED ? g * E.K : g;   // leaf slot of the slot's single evaluations (THREADED: worker 0's)
    bool threads_wait = false;               // THREADED: the workers of the search in progress hold the pending evaluations
    int search_root_vl = 0;                  // THREADED: root.virtual_loss left behind by the search that has just finished
    out_own = 0; out_opp = 0;
    if (pre.d.z) {   // leaf_waiting: this slot's leaf was emitted ahead of time (defer_leaf) and has not been evaluated yet
        out_own = E.leaf_own[ls]; out_opp = E.leaf_opp[ls];
        __syncwarp();
        if (lane == 0) {
            gp->leaf_waiting = 0;
            E.needs_eval[ls] = 1;
            if (count_tick) atomicAdd(E.ticks, 1ULL);
        }
        __syncwarp();
        return ADV_EMITTED;
    }
    // every lane keeps the (warp-uniform) scalar state in registers
    int phase = pre.a.x, sub_tree = pre.a.y, mover_tree = pre.a.z, sims_done = pre.a.w;
    int ply = pre.b.x, swap = pre.b.y, last_action = pre.b.z, n_moves_logged = pre.b.w;
    u64 game_index = pre.game_index;
    double my_noise = pre.noise;
    // the tree being worked on (the one the slot touches first: both came with the prefetch); the other one stays in gp->tree[]
    int curT = pre.c.x != PK_NONE ? pre.c.y : (phase == PH_SEARCH ? mover_tree : sub_tree);
    TreeState ts = curT ? pre.t1 : pre.t0;
    auto use = [&](const int T) {
        if (curT != T) {
            if (lane == 0) st_tree(&gp->tree[curT], ts);
            __syncwarp();
            ts = ld_tree(&gp->tree[T]);
            curT = T;
        }
        c.use_tree(T);
    };
#ifdef SPX_DBG_NO_COUNT   // timing experiment: no per-slot counters
    auto count = [&](u64*, const int) { };
#else
    auto count = [&](u64* counter, const int by) { if (lane == 0 && by) atomicAdd((ull*)counter, (ull)by); };   // RED.ADD: nothing to wait for
#endif
    bool emitted = false, parked = false, consumed = false;
    int out_net = 0;
    // The evaluation in hand: first the one the slot asked for on the previous tick (outputs my_p / v_in from the network), later
    // every request the evaluation cache answers on the spot; consumed at the top of the loop below.  When the slot emits, the
    // same variables describe the evaluation it will be waiting for (written to the slot in part 3).
    int e_kind = pre.c.x, e_tree = pre.c.y, e_parent = pre.c.z, e_action = pre.c.w, e_depth = pre.d.x, e_pplayer = pre.d.y;
    u64 e_own = pre.pend.x, e_opp = pre.pend.y;
    unsigned cp0 = pre.p0, cp1 = pre.p1;     // this lane's entries of that evaluation's select path
    float cur_p = my_p, cur_v = v_in;
    bool from_net = true;
    const bool use_cache = CACHE && !THREADED && E.ecache != nullptr;   // CACHE: only the fused tick kernel's instance carries the code
    if (THREADED && e_kind == PK_THREADS) { consumed = true; e_kind = PK_NONE; }   // the workers consume theirs inside the search round below

    // ---- run the state machine until the next network request: 1. consume the evaluation in hand, 2. go on
    while (!emitted) {
      if (e_kind != PK_NONE) {
        const int pend_kind = e_kind, T = e_tree, pend_parent = e_parent, pend_action = e_action, pend_depth = e_depth;
        const float my_p = cur_p, v_in = cur_v;
        consumed = true;
        e_kind = PK_NONE;
        use(T);
        if (use_cache && from_net) {   // remember what the network said about this position (net frame: mcts.py:316, modules.py:109-112)
            const bool root = pend_kind == PK_ROOT;
            ecache_insert<A>(E, g, lane, root ? 0ULL : (e_pplayer > 0 ? e_own : e_opp), root ? 0ULL : (e_pplayer > 0 ? e_opp : e_own),
                             cfg.two_nets ? T : 0, my_p, v_in);
        }
        if (pend_kind == PK_ROOT) {  // MCTreeSearch.reset (mcts.py:166-174); root.v is never read
            int player = (T == 0) ? (swap ? -1 : 1) : (swap ? 1 : -1);  // selfplayworker.py:175-176
            ts.node_count = 0;
            int idx = alloc_node<GAME>(c, ts, 0ULL, 0ULL, player, my_p);
            ts.root = idx; ts.root_n = 0; ts.root_w = 0.0; ts.root_player = player; ts.moves_played = 0; ts.n_rec = 0;
            if (THREADED && lane == 0) gp->root_vl[T] = 0;
            count(&gp->cnt_nodes, 1);
            if (T == 0) { n_moves_logged = 0; if (lane == 0) { E.own_action[2 * g] = 0; E.own_action[2 * g + 1] = -1; } }
            if (T == 0 && !cfg.opponent_kind) { phase = PH_RESET; sub_tree = 1; }
            else { mover_tree = swap ? 1 : 0; phase = PH_SEARCH; sims_done = -1; /* -1: search not begun */ }
        } else {
            // _expand_node's network branch (mcts.py:316-320) + backup (:361 / :207)
            const int pplayer = e_pplayer;
            int idx = alloc_node<GAME>(c, ts, e_own, e_opp, -pplayer, my_p);
            if (idx < 0) { count(&gp->cnt_err, 1); phase = PH_IDLE; }
            else {
                count(&gp->cnt_nodes, 1);
                if (lane == 0) ((int*)(c.node(pend_parent) + L::OFF_CHILD))[pend_action] = idx;
                double v = F::signed_by(F::from_f32(v_in), pplayer);  // modules.py:112 value*player
                unsigned p0 = 0, p1 = 0;
                if (pend_kind == PK_EXPAND) {
                    if (lane < pend_depth) p0 = cp0;
                    if (lane + 32 < pend_depth) p1 = cp1;
                } else if (lane == 0) p0 = ((unsigned)pend_parent << 4) | (unsigned)pend_action;
                __syncwarp();
                backup_path<GAME, SOFT>(c, p0, p1, pend_depth, v, ts);
                __syncwarp();
                if (pend_kind == PK_EXPAND) sims_done += 1;
                else {  // _set_root(node) (mcts.py:209)
                    if (THREADED && lane == 0) gp->root_vl[T] = (int)c.vbase[(size_t)pend_parent * L::VLS + pend_action];   // the child MCNode keeps its virtual_loss
                    ts.root = idx; ts.root_n = 1; ts.root_w = v; ts.root_player = -pplayer;
                    if (sub_tree == 0 && !cfg.opponent_kind) sub_tree = 1; else phase = PH_ENVSTEP;
                }
            }
        }
      }
        if (phase == PH_IDLE) break;
        if (phase == PH_RESET) {
            out_own = 0; out_opp = 0; out_net = cfg.two_nets ? sub_tree : 0;
            e_kind = PK_ROOT; e_tree = sub_tree;
            if (use_cache && ecache_lookup<A>(E, g, lane, 0ULL, 0ULL, out_net, cur_p, cur_v)) { from_net = false; out_own = 0; out_opp = 0; out_net = 0; count(&gp->cnt_hits, 1); continue; }
            emitted = true;
            break;
        }
        if (phase == PH_SEARCH && mover_tree == 1 && cfg.opponent_kind == SPX_OPP_EXTERNAL) {
            // the opposing player lives on the host (any BasePlayer): park until spx_set_external_actions delivers its move
            const int a = E.ext_action[g];
            if (a < 0) { parked = true; break; }
            __syncwarp();
            if (lane == 0) E.ext_action[g] = -1;
            last_action = a;
            phase = PH_REROOT; sub_tree = 0;
            continue;
        }
        if (phase == PH_SEARCH && mover_tree == 1 && cfg.opponent_kind) {
            // OneStepLookahead / Random (hardcoded_players.py:15-30,45-50).  The opponent's own env holds its pieces as +1:
            // own-frame "own" = env_opp, "enemy" = env_own; self.player = +1 if swap_sides else -1 (selfplayworker.py:176),
            // so without swap_sides the reference's "can I win" pass actually tests the ENEMY's move first -- kept as is.
            const int self_player = swap ? 1 : -1;
            const u64 env_own = gp->env_own, env_opp = gp->env_opp;
            const unsigned vmask = valid_mask<GAME>(env_own, env_opp);
            unsigned done_first = 0, done_second = 0;
            if (cfg.opponent_kind == SPX_OPP_LOOKAHEAD && lane < A && ((vmask >> lane) & 1u)) {
                for (int pass = 0; pass < 2; ++pass) {
                    u64 o = env_opp, e = env_own;   // opponent frame: own, enemy
                    int r = 0, dn = 0;
                    env_step<GAME>(o, e, lane, pass == 0 ? self_player : -self_player, r, dn);
                    if (dn) { if (pass == 0) done_first = 1; else done_second = 1; }
                }
            }
            const unsigned b0 = __ballot_sync(0xffffffffu, done_first != 0), b1 = __ballot_sync(0xffffffffu, done_second != 0);
            int action;
            if (b0) action = __ffs(b0) - 1;
            else if (b1) action = __ffs(b1) - 1;
            else {
                const int n = __popc(vmask);
                const double u = rng_uniform_from(rng_prefix(cfg.seed, game_index, 1, PURPOSE_OPPONENT, ply), 0, 0, 0);
                int idx = (int)(u * (double)n);
                if (idx >= n) idx = n - 1;
                unsigned m = vmask;
                for (int i = 0; i < idx; ++i) m &= m - 1;   // drop the idx lowest legal moves
                action = __ffs(m) - 1;
            }
            if (E.mlog && n_moves_logged < SPX_MAX_PLIES) {
                spx_move_log* ml = E.mlog + (size_t)g * SPX_MAX_PLIES + n_moves_logged;
                if (lane < A) { ml->n[lane] = 0; ml->w[lane] = 0.0; ml->noise[lane] = 0.0; }
                if (lane == 0) { ml->tree = 1; ml->ply = ply; ml->action = action; ml->root_n = 0; ml->root_w = 0.0; }
                n_moves_logged += 1;
            }
            last_action = action;
            phase = PH_REROOT; sub_tree = 0;
            continue;
        }
        if (phase == PH_SEARCH) {
            const int T = mover_tree;
            use(T);
            if (sims_done < 0) {  // search(): root.add_noise() (mcts.py:323-327, 49-53)
                double d = 1.0 / (double)A;
                if (cfg.noise_mode == 1 && E.noise_table) {
                    long long row = (long long)game_index - E.table_first;
                    int mv = ts.moves_played < E.table_moves ? ts.moves_played : E.table_moves - 1;
                    if (row >= 0 && row < E.table_games && lane < A)
                        d = E.noise_table[(((size_t)row * 2 + T) * E.table_moves + mv) * A + lane];
                } else if (cfg.noise_mode == 2) {
                    u64 pre_g = rng_prefix(cfg.seed, game_index, T, PURPOSE_GAMMA, ply);
                    double gv = lane < A ? gamma_variate(pre_g, lane, cfg.alpha) : 0.0;
                    double acc = 0.0;
                    for (int a = 0; a < A; ++a) acc = __dadd_rn(acc, shfl_d(gv, a));
                    d = __dmul_rn(gv, __ddiv_rn(1.0, acc));
                }
                if (lane < A) E.noise[(size_t)g * SPX_MAX_ACTIONS + lane] = d;
                my_noise = lane < A ? d : 0.0;
                sims_done = 0;
            }
            if constexpr (THREADED) {
                // thread_count = K: one round of the cooperative schedule per tick; sims_done counts the tasks STARTED
                int rvl = gp->root_vl[T];
                const bool finished = search_round_threaded<GAME>(E, c, gp, g, lane, T, ts, rvl, sims_done, game_index, ply, my_noise, policy_in, value_in);
                __syncwarp();
                if (lane == 0) gp->root_vl[T] = rvl;
                if (!finished) { threads_wait = true; e_tree = T; emitted = true; break; }
                search_root_vl = rvl;
            }
            if (sims_done >= cfg.sims) {
                // ---------------- _play (mcts.py:272-299)
                char* root = c.node(ts.root);
                int n_a = lane < A ? ((int*)(root + L::OFF_N))[lane] : 0;
                double w_a = lane < A ? ((double*)(root + L::OFF_W))[lane] : 0.0;
                double pw = (double)n_a;                                   // temp == 1 (mcts.py:182-183)
                if (cfg.evaluate) pw = lane < A ? pow_int_exact((unsigned)n_a, 20) : 0.0;  // temp/20 -> n**20.0
                double sum = 0.0;
                for (int a = 0; a < A; ++a) sum = __dadd_rn(sum, shfl_d(pw, a));
                int action = 0;
                const bool bad = !(sum > 0.0) || isinf(sum);  // NaN in p -> ValueError branch (:290-295)
                double prob = 0.0;
                if (bad) {
                    int best = __shfl_sync(0xffffffffu, n_a, 0);
                    for (int a = 1; a < A; ++a) { int na = __shfl_sync(0xffffffffu, n_a, a); if (na > best) { best = na; action = a; } }
                } else {
                    prob = __ddiv_rn(pw, sum);
                    double acc = 0.0;
                    for (int a = 0; a < A; ++a) acc = __dadd_rn(acc, shfl_d(prob, a));
                    const double last = acc;                                             // cdf[-1]
                    u64 pre_a = rng_prefix(cfg.seed, game_index, T, PURPOSE_ACTION, ply);
                    double u = rng_uniform_from(pre_a, 0, 0, 0);
                    acc = 0.0;
                    for (int a = 0; a < A; ++a) {   // searchsorted(cdf / cdf[-1], u, side="right") as np.random.choice does
                        acc = __dadd_rn(acc, shfl_d(prob, a));
                        if (__ddiv_rn(acc, last) <= u) action = a + 1;
                    }
                    if (action >= A) action = A - 1;
                    if (cfg.emit_records && ts.n_rec < SPX_MAX_OWN_MOVES) {
                        spx_record* rec = E.temp_rec + ((size_t)g * 2 + T) * SPX_MAX_OWN_MOVES + ts.n_rec;
                        if (lane < SPX_MAX_ACTIONS) rec->tree_probs[lane] = lane < A ? (float)prob : 0.f;
                        if (lane == 0) {
                            rec->own = *(u64*)(root + L::OFF_OWN);
                            rec->opp = *(u64*)(root + L::OFF_OPP);
                            rec->game_index = game_index;
                            // root.q (mcts.py:59-62): (w - virtual_loss) / (n + virtual_loss); the sequential search leaves virtual_loss = 0
                            const int q_den = ts.root_n + search_root_vl;
                            rec->q = q_den ? (float)__ddiv_rn(__dsub_rn(ts.root_w, (double)search_root_vl), (double)q_den) : 0.f;
                            rec->actual_val = 0.f;
                            rec->tree = (uint8_t)T; rec->ply = (uint8_t)ply; rec->pad0 = 0; rec->pad1 = 0;
                        }
                        ts.n_rec += 1;
                    }
                }
                if (E.mlog && n_moves_logged < SPX_MAX_PLIES) {
                    spx_move_log* ml = E.mlog + (size_t)g * SPX_MAX_PLIES + n_moves_logged;
                    if (lane < A) { ml->n[lane] = n_a; ml->w[lane] = w_a; ml->noise[lane] = my_noise; }
                    if (lane == 0) { ml->tree = T; ml->ply = ply; ml->action = action; ml->root_n = ts.root_n; ml->root_w = ts.root_w; }
                    n_moves_logged += 1;
                }
                ts.moves_played += 1;
                count(&gp->cnt_moves, 1);
                if (lane == 0 && T == 0) { E.own_action[2 * g] = ts.moves_played; E.own_action[2 * g + 1] = action; }
                last_action = action;
                phase = PH_REROOT; sub_tree = 0;
                continue;
            }
            if (budget <= 0) break;
            budget -= 1;
            // ---------------- search_node (mcts.py:340-367), sequential mode
            const u64 tie_pre = rng_prefix(cfg.seed, game_index, T, PURPOSE_TIE, ply);
            int node = ts.root, N = ts.root_n, player = ts.root_player, depth = 0;
            unsigned p0 = 0, p1 = 0;
            int child = 0, act = 0;
            for (;;) {
                const char* nd = c.node(node);
                double score = -INFINITY;
                int ch = 0, n = 0;
                double w = 0.0;
                float p = 0.f;
                unsigned meta = 0;
                if (lane < A) {
                    w = ((const double*)(nd + L::OFF_W))[lane];
                    n = ((const int*)(nd + L::OFF_N))[lane];
                    p = ((const float*)(nd + L::OFF_P))[lane];
                    ch = ((const int*)(nd + L::OFF_CHILD))[lane];
                    meta = *(const unsigned*)(nd + L::OFF_META);
                }
                // everything that does not depend on this node's statistics is computed while its loads are in flight:
                // sqrt(N + 1) (N came with the parent edge) and the tie-break noise of this (sim, depth, lane)
                const double sqrt_n = F::sqrt_int(N + 1, E.sqrt_table);
                const double tie = cfg.tie_mode ? F::mul(0.000001, F::from_u53(rng_uniform_bits(tie_pre, (unsigned)sims_done, (unsigned)depth, (u64)lane))) : 0.0;
                if (lane < A) {
                    if (ch >= 0) {   // pull every expanded child towards the SM while the scores are computed (the next level is one of them)
                        const char* cn = c.node(ch);
#ifndef SPX_DBG_NO_PREFETCH
                        asm volatile(SPX_PREFETCH " [%0];" ::"l"(cn));
                        asm volatile(SPX_PREFETCH " [%0];" ::"l"(cn + 128));
#endif
                    }
                    if ((meta >> lane) & 1u) {
                        const double q = n ? F::div_int(w, n) : 0.0;                                      // :59-62 (vl = 0)
                        double p_eff = F::from_f32(p);
                        if (depth == 0) p_eff = F::add(F::quarter(my_noise), F::mul(p_eff, 0.75));        // :64-69
                        const double u = F::div_int(F::mul(F::times4(p_eff), sqrt_n), 1 + n);             // :71-78
                        score = F::add(F::signed_by(q, player), u);                                       // :80-84
                    } else score = -10000000000.0;                                                        // :346-348
                    if (cfg.tie_mode) score = F::add(score, tie);
                }
                // np.argmax: first maximum wins (A <= 8: three butterfly rounds over 8 lanes suffice)
                int best = lane;
                double bs = score;
#pragma unroll
                for (int off = (A <= 8 ? 4 : 8); off > 0; off >>= 1) {
                    double os = __shfl_xor_sync(0xffffffffu, bs, off);
                    int ob = __shfl_xor_sync(0xffffffffu, best, off);
                    if (F::gt(os, bs) || (F::eq(os, bs) && ob < best)) { bs = os; best = ob; }
                }
                best = __shfl_sync(0xffffffffu, best, 0);
                child = __shfl_sync(0xffffffffu, ch, best);
                const int n_edge = __shfl_sync(0xffffffffu, n, best);
                const unsigned entry = ((unsigned)node << 4) | (unsigned)best;
                if (lane == (depth & 31)) { if (depth < 32) p0 = entry; else p1 = entry; }
                depth += 1;
                act = best;
                if (child < 0 || depth >= SPX_MAX_PATH) break;  // is_leaf: unexpanded or terminal (:357)
                node = child; N = n_edge; player = -player;
            }
            count(&gp->cnt_path, depth);
            // ---------------- _expand_node (mcts.py:301-321) on (node, act), mover = node.player
            const char* pnd = c.node(node);
            const u64 par_own = *(const u64*)(pnd + L::OFF_OWN), par_opp = *(const u64*)(pnd + L::OFF_OPP);
            u64 c_own = par_own, c_opp = par_opp;
            int r = 0, done = 0;
            if (child == CHILD_UNEXPANDED) env_step<GAME>(c_own, c_opp, act, player, r, done);
            else { done = 1; r = (child == CHILD_TERM_WIN); }
            count(&gp->cnt_sims, 1);   // (a sim that needs the network completes when its evaluation is consumed next tick)
            if (done) {
                if (child == CHILD_UNEXPANDED && lane == 0)
                    ((int*)(c.node(node) + L::OFF_CHILD))[act] = r ? CHILD_TERM_WIN : CHILD_TERM_DRAW;
                const double v = terminal_value(cfg.strong_play, r * player, par_own, par_opp);
                __syncwarp();
                backup_path<GAME, SOFT>(c, p0, p1, depth, v, ts);
                __syncwarp();
                sims_done += 1;
                count(&gp->cnt_term, 1);
                continue;
            }
            // needs the network: net input = child_state * parent.player (mcts.py:316, modules.py:109-112)
            out_own = player > 0 ? c_own : c_opp;
            out_opp = player > 0 ? c_opp : c_own;
            out_net = cfg.two_nets ? T : 0;
            unsigned* pp = E.paths + (size_t)g * SPX_MAX_PATH;
            if (lane < depth) pp[lane] = p0;
            if (lane + 32 < depth) pp[lane + 32] = p1;
            e_kind = PK_EXPAND; e_tree = T; e_parent = node; e_action = act; e_depth = depth;
            e_pplayer = player; e_own = c_own; e_opp = c_opp;
            if (use_cache && ecache_lookup<A>(E, g, lane, out_own, out_opp, out_net, cur_p, cur_v)) {
                from_net = false; cp0 = p0; cp1 = p1; out_own = 0; out_opp = 0; out_net = 0; count(&gp->cnt_hits, 1);
                continue;   // consumed at the top of the loop: create_children, backup, the next simulation
            }
            emitted = true;
            break;
        }
        if (phase == PH_REROOT) {
            // play_action -> _set_node (mcts.py:188-209) on tree sub_tree with last_action
            const int T = sub_tree, a = last_action;
            use(T);
            bool wait_net = false;
            if (ts.root < 0) count(&gp->cnt_err, 1);  // re-rooting a finished tree: cannot happen in legal play
            else {
                char* root = c.node(ts.root);
                const int n_a = ((const int*)(root + L::OFF_N))[a];
                const int ch = ((const int*)(root + L::OFF_CHILD))[a];
                if (n_a == 0) {
                    const u64 par_own = *(const u64*)(root + L::OFF_OWN), par_opp = *(const u64*)(root + L::OFF_OPP);
                    u64 c_own = par_own, c_opp = par_opp;
                    int r = 0, done = 0;
                    const int player = ts.root_player;
                    env_step<GAME>(c_own, c_opp, a, player, r, done);
                    if (done) {
                        const double v = terminal_value(cfg.strong_play, r * player, par_own, par_opp);
                        __syncwarp();
                        if (lane == 0) {
                            ((int*)(root + L::OFF_CHILD))[a] = r ? CHILD_TERM_WIN : CHILD_TERM_DRAW;
                            ((int*)(root + L::OFF_N))[a] = 1;
                            ((double*)(root + L::OFF_W))[a] = v;
                        }
                        __syncwarp();
                        ts.root = -1; ts.root_n = 1; ts.root_w = v; ts.root_player = -player;
                    } else {
                        out_own = player > 0 ? c_own : c_opp;
                        out_opp = player > 0 ? c_opp : c_own;
                        out_net = cfg.two_nets ? T : 0;
                        e_kind = PK_REROOT; e_tree = T; e_parent = ts.root; e_action = a; e_depth = 1;
                        e_pplayer = player; e_own = c_own; e_opp = c_opp;
                        wait_net = true;
                    }
                } else {
                    ts.root_w = ((const double*)(root + L::OFF_W))[a];
                    ts.root_n = n_a;
                    ts.root_player = -ts.root_player;
                    if (THREADED && lane == 0) gp->root_vl[T] = (int)c.vbase[(size_t)ts.root * L::VLS + a];   // the child MCNode keeps its virtual_loss
                    ts.root = ch >= 0 ? ch : -1;
                }
            }
            if (wait_net) {
                if (use_cache && ecache_lookup<A>(E, g, lane, out_own, out_opp, out_net, cur_p, cur_v)) { from_net = false; out_own = 0; out_opp = 0; out_net = 0; count(&gp->cnt_hits, 1); continue; }
                emitted = true;
                break;
            }
            if (T == 0 && !cfg.opponent_kind) sub_tree = 1; else phase = PH_ENVSTEP;
            continue;
        }
        if (phase == PH_ENVSTEP) {
            // env.step(a, player) in play_move (selfplayworker.py:221-224) and the episode bookkeeping (:180-190)
            const int player = mover_tree == 0 ? 1 : -1;
            u64 env_own = gp->env_own, env_opp = gp->env_opp;
            int r = 0, done = 0;
            const int st = env_step<GAME>(env_own, env_opp, last_action, player, r, done);
            if (st != SPX_ENV_OK) { count(&gp->cnt_err, 1); done = 1; }
            ply += 1;
            if (!done) {
                __syncwarp();
                if (lane == 0) { gp->env_own = env_own; gp->env_opp = env_opp; }
                __syncwarp();
                mover_tree ^= 1;
                phase = PH_SEARCH; sims_done = -1;
                continue;
            }
            const int reward = r * player;  // get_and_play_moves: r = r * player (:218)
            if (lane == 0) {
                unsigned long long k = atomicAdd(E.res_count, 1ULL);
                if (k < (unsigned long long)cfg.result_capacity) {
                    spx_result res;
                    memset(&res, 0, sizeof(res));
                    res.game_index = game_index; res.reward = (int8_t)reward; res.swap_sides = (uint8_t)swap; res.plies = (uint8_t)ply;
                    E.res_ring[k] = res;
                }
            }
            // both trees' record counts: the current tree goes back to the slot first
            if (lane == 0) st_tree(&gp->tree[curT], ts);
            __syncwarp();
            const int n0 = gp->tree[0].n_rec, n1 = gp->tree[1].n_rec;
            if (cfg.emit_records) {  // push_to_queue (mcts.py:225-232): policy first (+r), then the opponent (-r)
                const int tot = n0 + n1;
                unsigned long long base = 0;
                if (lane == 0) base = atomicAdd(E.rec_count, (unsigned long long)tot);
                base = __shfl_sync(0xffffffffu, base, 0);
                for (int i = lane; i < tot; i += 32) {
                    const int T = i < n0 ? 0 : 1, j = i < n0 ? i : i - n0;
                    spx_record rec = E.temp_rec[((size_t)g * 2 + T) * SPX_MAX_OWN_MOVES + j];
                    rec.actual_val = (float)(T == 0 ? reward : -reward);
                    if (base + i < (unsigned long long)cfg.record_capacity) E.rec_ring[base + i] = rec;
                    else atomicAdd(E.rec_dropped, 1ULL);
                }
            }
            count(&gp->cnt_games, 1);
            // next game on this slot (self_play_parallel.py:250-253: swap_sides = game index odd)
            game_index += (u64)cfg.slot_stride;
            ply = 0;
            swap = (int)(game_index & 1ULL);
            __syncwarp();
            if (lane == 0) { gp->env_own = 0; gp->env_opp = 0; gp->tree[0].n_rec = 0; gp->tree[1].n_rec = 0; }
            __syncwarp();
            ts.n_rec = 0;
            if ((long long)game_index < cfg.games_target) { phase = PH_RESET; sub_tree = 0; }
            else phase = PH_IDLE;
            continue;
        }
        break;
    }

    // ---- 3. write the slot back: 128-bit stores by lane 0, nothing the warp has to wait for
    __syncwarp();
    if (lane == 0) {
        int4* q = reinterpret_cast<int4*>(gp);
        q[0] = make_int4(phase, sub_tree, mover_tree, sims_done);
        q[1] = make_int4(ply, swap, last_action, n_moves_logged);
        if (threads_wait) q[2] = make_int4(PK_THREADS, e_tree, 0, 0);    // the leaf slots were written by the search round
        else if (emitted) {
            q[2] = make_int4(e_kind, e_tree, e_parent, e_action);
            q[3] = make_int4(e_depth, e_pplayer, defer_leaf ? 1 : 0, 0);
            reinterpret_cast<ulonglong2*>(gp)[4] = make_ulonglong2(e_own, e_opp);
            atomicAdd((ull*)&gp->cnt_evals, 1ULL);
        } else if (consumed) q[2] = make_int4(PK_NONE, 0, 0, 0);
        gp->game_index = game_index;
        st_tree(&gp->tree[curT], ts);
        if (!threads_wait) {
            E.leaf_own[ls] = out_own;
            E.leaf_opp[ls] = out_opp;
            E.needs_eval[ls] = (emitted && !defer_leaf) ? 1 : 0;
            E.net_id[ls] = (unsigned char)out_net;
        }
        if (count_tick) atomicAdd(E.ticks, 1ULL);
    }
    if (THREADED && !threads_wait && lane >= 1 && lane < E.K) E.needs_eval[ls + lane] = 0;   // single evaluations use worker 0's slot
    __syncwarp();   // lane 0's stores are ordered before whatever any lane of this warp loads next (the next prefetch of this slot)
    return emitted ? ADV_EMITTED : (phase == PH_IDLE ? ADV_IDLE : (parked ? ADV_PARKED : 0));
}

}  // namespace spx

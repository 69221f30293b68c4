// spx_advance.cuh -- the search engine's device side: node-pool layout, per-slot state, and advance_game(), the per-warp state
// machine of one game slot (see the header comment of spx_engine.cu).  Shared by spx_engine.cu (advance_kernel: one warp per
// game) and spx_tower.cu (fused tick kernel: the network CTA that owns a game's leaf also advances the game).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <string.h>

#include "spx_common.cuh"

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
    static_assert(OFF_OWN % 8 == 0, "bitboards must be 8-byte aligned");
};

enum { CHILD_UNEXPANDED = -1, CHILD_TERM_DRAW = -2, CHILD_TERM_WIN = -3 };
enum { PH_IDLE = 0, PH_RESET = 1, PH_SEARCH = 2, PH_REROOT = 3, PH_ENVSTEP = 4 };
enum { PK_NONE = 0, PK_ROOT = 1, PK_EXPAND = 2, PK_REROOT = 3 };
#define SPX_MAX_PATH 64
#define SPX_MAX_OWN_MOVES 22
#define SPX_MAX_PLIES 44

struct TreeState {
    double root_w;
    int root, root_n, root_player, moves_played, node_count, n_rec;
};

struct GameState {
    u64 env_own, env_opp;  // env frame: own = +1 = the policy (tree 0)
    u64 game_index;
    u64 pend_own, pend_opp;  // child state awaiting its evaluation, TREE frame
    TreeState tree[2];
    int phase, sub_tree, mover_tree, sims_done, ply, swap, last_action;
    int pend_kind, pend_tree, pend_parent, pend_action, pend_depth, pend_parent_player;
    int n_moves_logged, pad;
    u64 cnt_sims, cnt_evals, cnt_term, cnt_path, cnt_moves, cnt_games, cnt_nodes, cnt_err;
};

struct EngineDev {
    spx_config cfg;
    int nodes_per_tree;
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
};

}  // namespace spx

struct spx_engine {
    spx::EngineDev d;
    int device;
    int64_t bytes;
};

namespace spx {

// ------------------------------------------------------------------------------------------------ device helpers
__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }

// n**k, correctly rounded (twin of ox_pow_int_exact): np.power(n, 1/temp) for integral 1/temp (mcts.py:100-101)
__device__ inline double pow_int_exact(unsigned n, int k) {
    if (k == 0) return 1.0;
    if (n == 0) return 0.0;
    unsigned limb[40];
    int nl = 1;
    limb[0] = 1;
    for (int i = 0; i < k; ++i) {
        u64 carry = 0;
        for (int j = 0; j < nl; ++j) { u64 t = (u64)limb[j] * n + carry; limb[j] = (unsigned)t; carry = t >> 32; }
        if (carry) { if (nl >= 40) return INFINITY; limb[nl++] = (unsigned)carry; }
    }
    int top = nl - 1;
    while (top > 0 && limb[top] == 0) --top;
    int hb = 31;
    while (!((limb[top] >> hb) & 1)) --hb;
    int nbits = top * 32 + hb + 1;
    if (nbits <= 53) { double v = 0; for (int j = top; j >= 0; --j) v = __dadd_rn(__dmul_rn(v, 4294967296.0), (double)limb[j]); return v; }
    int shift = nbits - 53;
    u64 mant = 0;
    for (int b = nbits - 1; b >= shift; --b) mant = (mant << 1) | ((limb[b / 32] >> (b % 32)) & 1u);
    int half = (limb[(shift - 1) / 32] >> ((shift - 1) % 32)) & 1u;
    int sticky = 0;
    for (int b = shift - 2; b >= 0 && !sticky; --b) sticky |= (limb[b / 32] >> (b % 32)) & 1u;
    if (half && (sticky || (mant & 1))) mant += 1;
    return ldexp((double)mant, shift);
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
        double X = sqrt(-2.0 * log(1.0 - SPX_U(0))) * cos(6.283185307179586 * SPX_U(1));
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
#define SPX_PREFETCH "prefetch.global.L2"
#endif
template <int GAME> struct Ctx {
    typedef Rules<GAME> R;
    typedef NodeLayout<GAME> L;
    const EngineDev& E;
    int g, lane;
    char* tbase;  // node pool of the tree currently worked on
    __device__ Ctx(const EngineDev& e, int g_, int lane_) : E(e), g(g_), lane(lane_), tbase(nullptr) {}
    __device__ __forceinline__ void use_tree(int t) {
        tbase = E.pool + ((size_t)g * 2 + t) * (size_t)E.nodes_per_tree * L::SIZE;
    }
    __device__ __forceinline__ char* node(int idx) const { return tbase + (size_t)idx * L::SIZE; }
};

// terminal value v = r (already multiplied by the mover) or the strong_play formula (mcts.py:305-313)
__device__ __forceinline__ double terminal_value(int strong, int r_signed, u64 parent_own, u64 parent_opp) {
    if (!strong) return (double)r_signed;
    int num_steps = __popcll(parent_own | parent_opp) + 1;
    return __dmul_rn(__dsub_rn(1.18, __ddiv_rn((double)(9 * num_steps), 350.0)), (double)r_signed);
}

// w += v; n += 1 on every edge of the path and on the root itself (MCNode.backup, mcts.py:94-98;
// ancestors above the current root are never read again so they are skipped).
template <int GAME>
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
            *pw = __dadd_rn(*pw, v);
        }
    }
    ts.root_n += 1;
    ts.root_w = __dadd_rn(ts.root_w, v);
}

// create_children (mcts.py:103-107) for a freshly evaluated position + link from its parent edge
template <int GAME>
__device__ __forceinline__ int alloc_node(const Ctx<GAME>& c, TreeState& ts, u64 own, u64 opp, int player, const float* policy) {
    typedef NodeLayout<GAME> L;
    typedef Rules<GAME> R;
    int idx = ts.node_count;
    if (idx >= c.E.nodes_per_tree) return -1;
    ts.node_count = idx + 1;
    char* nd = c.node(idx);
    if (c.lane < R::A) {
        ((double*)(nd + L::OFF_W))[c.lane] = 0.0;
        ((int*)(nd + L::OFF_N))[c.lane] = 0;
        ((float*)(nd + L::OFF_P))[c.lane] = policy[c.lane];
        ((int*)(nd + L::OFF_CHILD))[c.lane] = CHILD_UNEXPANDED;
    }
    if (c.lane == 0) {
        *(unsigned*)(nd + L::OFF_META) = valid_mask<GAME>(own, opp) | (player < 0 ? 0x10000u : 0u);
        *(u64*)(nd + L::OFF_OWN) = own;
        *(u64*)(nd + L::OFF_OPP) = opp;
    }
    return idx;
}

// ------------------------------------------------------------------------------------------------ one game, one tick
// One game slot's share of a tick, executed by one whole warp (all 32 lanes call it together): consume the evaluation the slot
// asked for, run its state machine until the next network request, publish the leaf.  Called by advance_kernel (one warp per
// game) and by the fused tick kernel of spx_tower.cu (the epilogue warps of the CTA that evaluates the game's leaf).
template <int GAME>
// (policy_in / value_in are deliberately not __restrict__: in the fused kernel they were written earlier in the same launch.)
__device__ __forceinline__ void advance_game(const EngineDev& E, const int g, const int lane, const float* policy_in, const float* value_in) {
    typedef Rules<GAME> R;
    typedef NodeLayout<GAME> L;
    constexpr int A = R::A;
    Ctx<GAME> c(E, g, lane);
    GameState* gp = E.games + g;
    GameState s = *gp;  // every lane keeps the (warp-uniform) scalar state in registers
    const spx_config& cfg = E.cfg;

    bool emitted = false;
    u64 out_own = 0, out_opp = 0;
    int out_net = 0;

    // ---- 1. consume the evaluation this slot asked for on the previous tick
    if (s.pend_kind != PK_NONE) {
        const float* pol = policy_in + (size_t)g * A;
        const int T = s.pend_tree;
        c.use_tree(T);
        TreeState ts = s.tree[T];
        if (s.pend_kind == PK_ROOT) {  // MCTreeSearch.reset (mcts.py:166-174); root.v is never read
            int player = (T == 0) ? (s.swap ? -1 : 1) : (s.swap ? 1 : -1);  // selfplayworker.py:175-176
            ts.node_count = 0;
            int idx = alloc_node<GAME>(c, ts, 0ULL, 0ULL, player, pol);
            ts.root = idx; ts.root_n = 0; ts.root_w = 0.0; ts.root_player = player; ts.moves_played = 0; ts.n_rec = 0;
            s.cnt_nodes += 1;
            s.tree[T] = ts;
            if (T == 0) { s.n_moves_logged = 0; if (lane == 0) { E.own_action[2 * g] = 0; E.own_action[2 * g + 1] = -1; } }
            if (T == 0 && !cfg.opponent_kind) { s.phase = PH_RESET; s.sub_tree = 1; }
            else { s.mover_tree = s.swap ? 1 : 0; s.phase = PH_SEARCH; s.sims_done = -1; /* -1: search not begun */ }
        } else {
            // _expand_node's network branch (mcts.py:316-320) + backup (:361 / :207)
            const int pplayer = s.pend_parent_player;
            int idx = alloc_node<GAME>(c, ts, s.pend_own, s.pend_opp, -pplayer, pol);
            if (idx < 0) { s.cnt_err += 1; s.phase = PH_IDLE; }
            else {
                s.cnt_nodes += 1;
                if (lane == 0) ((int*)(c.node(s.pend_parent) + L::OFF_CHILD))[s.pend_action] = idx;
                double v = __dmul_rn((double)value_in[g], (double)pplayer);  // modules.py:112 value*player
                unsigned p0 = 0, p1 = 0;
                if (s.pend_kind == PK_EXPAND) {
                    const unsigned* pp = E.paths + (size_t)g * SPX_MAX_PATH;
                    if (lane < s.pend_depth) p0 = pp[lane];
                    if (lane + 32 < s.pend_depth) p1 = pp[lane + 32];
                } else if (lane == 0) p0 = ((unsigned)s.pend_parent << 4) | (unsigned)s.pend_action;
                __syncwarp();
                backup_path<GAME>(c, p0, p1, s.pend_depth, v, ts);
                __syncwarp();
                if (s.pend_kind == PK_EXPAND) s.sims_done += 1;
                else {  // _set_root(node) (mcts.py:209)
                    ts.root = idx; ts.root_n = 1; ts.root_w = v; ts.root_player = -pplayer;
                    if (s.sub_tree == 0 && !cfg.opponent_kind) s.sub_tree = 1; else s.phase = PH_ENVSTEP;
                }
            }
            s.tree[T] = ts;
        }
        s.pend_kind = PK_NONE;
    }

    // ---- 2. run the state machine until the next network request
    int budget = cfg.max_sims_per_tick;
    while (!emitted) {
        if (s.phase == PH_IDLE) break;
        if (s.phase == PH_RESET) {
            out_own = 0; out_opp = 0; out_net = cfg.two_nets ? s.sub_tree : 0;
            s.pend_kind = PK_ROOT; s.pend_tree = s.sub_tree;
            emitted = true;
            break;
        }
        if (s.phase == PH_SEARCH && s.mover_tree == 1 && cfg.opponent_kind == SPX_OPP_EXTERNAL) {
            // the opposing player lives on the host (any BasePlayer): park until spx_set_external_actions delivers its move
            const int a = E.ext_action[g];
            if (a < 0) break;
            __syncwarp();
            if (lane == 0) E.ext_action[g] = -1;
            s.last_action = a;
            s.phase = PH_REROOT; s.sub_tree = 0;
            continue;
        }
        if (s.phase == PH_SEARCH && s.mover_tree == 1 && cfg.opponent_kind) {
            // OneStepLookahead / Random (hardcoded_players.py:15-30,45-50).  The opponent's own env holds its pieces as +1:
            // own-frame "own" = env_opp, "enemy" = env_own; self.player = +1 if swap_sides else -1 (selfplayworker.py:176),
            // so without swap_sides the reference's "can I win" pass actually tests the ENEMY's move first -- kept as is.
            const int self_player = s.swap ? 1 : -1;
            const unsigned vmask = valid_mask<GAME>(s.env_own, s.env_opp);
            unsigned done_first = 0, done_second = 0;
            if (cfg.opponent_kind == SPX_OPP_LOOKAHEAD && lane < A && ((vmask >> lane) & 1u)) {
                for (int pass = 0; pass < 2; ++pass) {
                    u64 o = s.env_opp, e = s.env_own;   // opponent frame: own, enemy
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
                const double u = rng_uniform_from(rng_prefix(cfg.seed, s.game_index, 1, PURPOSE_OPPONENT, s.ply), 0, 0, 0);
                int idx = (int)(u * (double)n);
                if (idx >= n) idx = n - 1;
                unsigned m = vmask;
                for (int i = 0; i < idx; ++i) m &= m - 1;   // drop the idx lowest legal moves
                action = __ffs(m) - 1;
            }
            if (E.mlog && s.n_moves_logged < SPX_MAX_PLIES) {
                spx_move_log* ml = E.mlog + (size_t)g * SPX_MAX_PLIES + s.n_moves_logged;
                if (lane < A) { ml->n[lane] = 0; ml->w[lane] = 0.0; ml->noise[lane] = 0.0; }
                if (lane == 0) { ml->tree = 1; ml->ply = s.ply; ml->action = action; ml->root_n = 0; ml->root_w = 0.0; }
                s.n_moves_logged += 1;
            }
            s.last_action = action;
            s.phase = PH_REROOT; s.sub_tree = 0;
            continue;
        }
        if (s.phase == PH_SEARCH) {
            const int T = s.mover_tree;
            c.use_tree(T);
            TreeState ts = s.tree[T];
            if (s.sims_done < 0) {  // search(): root.add_noise() (mcts.py:323-327, 49-53)
                double d = 1.0 / (double)A;
                if (cfg.noise_mode == 1 && E.noise_table) {
                    long long row = (long long)s.game_index - E.table_first;
                    int mv = ts.moves_played < E.table_moves ? ts.moves_played : E.table_moves - 1;
                    if (row >= 0 && row < E.table_games && lane < A)
                        d = E.noise_table[(((size_t)row * 2 + T) * E.table_moves + mv) * A + lane];
                } else if (cfg.noise_mode == 2) {
                    u64 pre = rng_prefix(cfg.seed, s.game_index, T, PURPOSE_GAMMA, s.ply);
                    double gv = lane < A ? gamma_variate(pre, lane, cfg.alpha) : 0.0;
                    double acc = 0.0;
                    for (int a = 0; a < A; ++a) acc = __dadd_rn(acc, shfl_d(gv, a));
                    d = __dmul_rn(gv, __ddiv_rn(1.0, acc));
                }
                if (lane < A) E.noise[(size_t)g * SPX_MAX_ACTIONS + lane] = d;
                __syncwarp();
                s.sims_done = 0;
            }
            if (s.sims_done >= cfg.sims) {
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
                    double cdf[SPX_MAX_ACTIONS], acc = 0.0;
                    for (int a = 0; a < A; ++a) { acc = __dadd_rn(acc, shfl_d(prob, a)); cdf[a] = acc; }
                    u64 pre = rng_prefix(cfg.seed, s.game_index, T, PURPOSE_ACTION, s.ply);
                    double u = rng_uniform_from(pre, 0, 0, 0);
                    for (int a = 0; a < A; ++a) if (__ddiv_rn(cdf[a], cdf[A - 1]) <= u) action = a + 1;  // searchsorted right
                    if (action >= A) action = A - 1;
                    if (cfg.emit_records && ts.n_rec < SPX_MAX_OWN_MOVES) {
                        spx_record* rec = E.temp_rec + ((size_t)g * 2 + T) * SPX_MAX_OWN_MOVES + ts.n_rec;
                        if (lane < SPX_MAX_ACTIONS) rec->tree_probs[lane] = lane < A ? (float)prob : 0.f;
                        if (lane == 0) {
                            rec->own = *(u64*)(root + L::OFF_OWN);
                            rec->opp = *(u64*)(root + L::OFF_OPP);
                            rec->game_index = s.game_index;
                            rec->q = ts.root_n ? (float)__ddiv_rn(ts.root_w, (double)ts.root_n) : 0.f;  // root.q
                            rec->actual_val = 0.f;
                            rec->tree = (uint8_t)T; rec->ply = (uint8_t)s.ply; rec->pad0 = 0; rec->pad1 = 0;
                        }
                        ts.n_rec += 1;
                    }
                }
                if (E.mlog && s.n_moves_logged < SPX_MAX_PLIES) {
                    spx_move_log* ml = E.mlog + (size_t)g * SPX_MAX_PLIES + s.n_moves_logged;
                    if (lane < A) { ml->n[lane] = n_a; ml->w[lane] = w_a; ml->noise[lane] = E.noise[(size_t)g * SPX_MAX_ACTIONS + lane]; }
                    if (lane == 0) { ml->tree = T; ml->ply = s.ply; ml->action = action; ml->root_n = ts.root_n; ml->root_w = ts.root_w; }
                    s.n_moves_logged += 1;
                }
                ts.moves_played += 1;
                s.cnt_moves += 1;
                if (lane == 0 && T == 0) { E.own_action[2 * g] = ts.moves_played; E.own_action[2 * g + 1] = action; }
                s.tree[T] = ts;
                s.last_action = action;
                s.phase = PH_REROOT; s.sub_tree = 0;
                continue;
            }
            if (budget <= 0) break;
            budget -= 1;
            // ---------------- search_node (mcts.py:340-367), sequential mode
            const u64 tie_pre = rng_prefix(cfg.seed, s.game_index, T, PURPOSE_TIE, s.ply);
            const double my_noise = lane < A ? E.noise[(size_t)g * SPX_MAX_ACTIONS + lane] : 0.0;
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
                const double sqrt_n = __dsqrt_rn((double)(N + 1));
                const double tie = cfg.tie_mode ? __dmul_rn(0.000001, rng_uniform_from(tie_pre, (unsigned)s.sims_done, (unsigned)depth, (u64)lane)) : 0.0;
                if (lane < A) {
                    if (ch >= 0) {   // pull every expanded child towards the SM while the scores are computed (the next level is one of them)
                        const char* cn = c.node(ch);
                        asm volatile(SPX_PREFETCH " [%0];" ::"l"(cn));
                        asm volatile(SPX_PREFETCH " [%0];" ::"l"(cn + 128));
                    }
                    if ((meta >> lane) & 1u) {
                        const double q = n ? __ddiv_rn(w, (double)n) : 0.0;                               // :59-62 (vl = 0)
                        double p_eff = (double)p;
                        if (depth == 0) p_eff = __dadd_rn(__dmul_rn(my_noise, 0.25), __dmul_rn((double)p, 0.75));  // :64-69
                        const double u = __ddiv_rn(__dmul_rn(__dmul_rn(4.0, p_eff), sqrt_n), (double)(1 + n));   // :71-78
                        score = __dadd_rn(__dmul_rn((double)player, q), u);                               // :80-84
                    } else score = -10000000000.0;                                                        // :346-348
                    if (cfg.tie_mode) score = __dadd_rn(score, tie);
                }
                // np.argmax: first maximum wins (A <= 8: three butterfly rounds over 8 lanes suffice)
                int best = lane;
                double bs = score;
#pragma unroll
                for (int off = (A <= 8 ? 4 : 8); off > 0; off >>= 1) {
                    double os = __shfl_xor_sync(0xffffffffu, bs, off);
                    int ob = __shfl_xor_sync(0xffffffffu, best, off);
                    if (os > bs || (os == bs && ob < best)) { bs = os; best = ob; }
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
            s.cnt_path += (u64)depth;
            // ---------------- _expand_node (mcts.py:301-321) on (node, act), mover = node.player
            const char* pnd = c.node(node);
            const u64 par_own = *(const u64*)(pnd + L::OFF_OWN), par_opp = *(const u64*)(pnd + L::OFF_OPP);
            u64 c_own = par_own, c_opp = par_opp;
            int r = 0, done = 0;
            if (child == CHILD_UNEXPANDED) env_step<GAME>(c_own, c_opp, act, player, r, done);
            else { done = 1; r = (child == CHILD_TERM_WIN); }
            if (done) {
                if (child == CHILD_UNEXPANDED && lane == 0)
                    ((int*)(c.node(node) + L::OFF_CHILD))[act] = r ? CHILD_TERM_WIN : CHILD_TERM_DRAW;
                const double v = terminal_value(cfg.strong_play, r * player, par_own, par_opp);
                __syncwarp();
                backup_path<GAME>(c, p0, p1, depth, v, ts);
                __syncwarp();
                s.sims_done += 1; s.cnt_sims += 1; s.cnt_term += 1;
                s.tree[T] = ts;
                continue;
            }
            // needs the network: net input = child_state * parent.player (mcts.py:316, modules.py:109-112)
            out_own = player > 0 ? c_own : c_opp;
            out_opp = player > 0 ? c_opp : c_own;
            out_net = cfg.two_nets ? T : 0;
            unsigned* pp = E.paths + (size_t)g * SPX_MAX_PATH;
            if (lane < depth) pp[lane] = p0;
            if (lane + 32 < depth) pp[lane + 32] = p1;
            s.pend_kind = PK_EXPAND; s.pend_tree = T; s.pend_parent = node; s.pend_action = act; s.pend_depth = depth;
            s.pend_parent_player = player; s.pend_own = c_own; s.pend_opp = c_opp;
            s.cnt_sims += 1;  // the sim completes when its evaluation is consumed next tick
            s.tree[T] = ts;
            emitted = true;
            break;
        }
        if (s.phase == PH_REROOT) {
            // play_action -> _set_node (mcts.py:188-209) on tree sub_tree with s.last_action
            const int T = s.sub_tree, a = s.last_action;
            c.use_tree(T);
            TreeState ts = s.tree[T];
            bool parked = false;
            if (ts.root < 0) s.cnt_err += 1;  // re-rooting a finished tree: cannot happen in legal play
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
                        s.pend_kind = PK_REROOT; s.pend_tree = T; s.pend_parent = ts.root; s.pend_action = a; s.pend_depth = 1;
                        s.pend_parent_player = player; s.pend_own = c_own; s.pend_opp = c_opp;
                        parked = true;
                    }
                } else {
                    ts.root_w = ((const double*)(root + L::OFF_W))[a];
                    ts.root_n = n_a;
                    ts.root_player = -ts.root_player;
                    ts.root = ch >= 0 ? ch : -1;
                }
            }
            s.tree[T] = ts;
            if (parked) { emitted = true; break; }
            if (T == 0 && !cfg.opponent_kind) s.sub_tree = 1; else s.phase = PH_ENVSTEP;
            continue;
        }
        if (s.phase == PH_ENVSTEP) {
            // env.step(a, player) in play_move (selfplayworker.py:221-224) and the episode bookkeeping (:180-190)
            const int player = s.mover_tree == 0 ? 1 : -1;
            int r = 0, done = 0;
            const int st = env_step<GAME>(s.env_own, s.env_opp, s.last_action, player, r, done);
            if (st != SPX_ENV_OK) { s.cnt_err += 1; done = 1; }
            s.ply += 1;
            if (!done) {
                s.mover_tree ^= 1;
                s.phase = PH_SEARCH; s.sims_done = -1;
                continue;
            }
            const int reward = r * player;  // get_and_play_moves: r = r * player (:218)
            if (lane == 0) {
                unsigned long long k = atomicAdd(E.res_count, 1ULL);
                if (k < (unsigned long long)cfg.result_capacity) {
                    spx_result res;
                    memset(&res, 0, sizeof(res));
                    res.game_index = s.game_index; res.reward = (int8_t)reward; res.swap_sides = (uint8_t)s.swap; res.plies = (uint8_t)s.ply;
                    E.res_ring[k] = res;
                }
            }
            if (cfg.emit_records) {  // push_to_queue (mcts.py:225-232): policy first (+r), then the opponent (-r)
                const int n0 = s.tree[0].n_rec, n1 = s.tree[1].n_rec, tot = n0 + n1;
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
            s.cnt_games += 1;
            // next game on this slot (self_play_parallel.py:250-253: swap_sides = game index odd)
            s.game_index += (u64)cfg.slot_stride;
            s.env_own = 0; s.env_opp = 0; s.ply = 0;
            s.swap = (int)(s.game_index & 1ULL);
            s.tree[0].n_rec = 0; s.tree[1].n_rec = 0;
            if ((long long)s.game_index < cfg.games_target) { s.phase = PH_RESET; s.sub_tree = 0; }
            else s.phase = PH_IDLE;
            continue;
        }
        break;
    }

    if (emitted) s.cnt_evals += 1;
    __syncwarp();
    if (lane == 0) {
        *gp = s;
        E.leaf_own[g] = out_own;
        E.leaf_opp[g] = out_opp;
        E.needs_eval[g] = emitted ? 1 : 0;
        E.net_id[g] = (unsigned char)out_net;
        if (g == 0) *E.ticks += 1ULL;
    }
}

}  // namespace spx

/*
 * spx_oracle.c -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product path (libspx.so, CUDA) never links or calls it.
 *
 * It restates, in the reference's own direct (recursive, array-board) style:
 *   games/connect4/connect4env.py:29-95      Connect4Env.step/get_reward/valid_moves/set_state
 *   games/tictactoe/tictactoe_env.py:23-82   TicTacToeEnv.step/get_reward/valid_moves
 *   games/algos/mcts.py:21-113               MCNode (q, p_eff, u, select_prob, backup, ...)
 *   games/algos/mcts.py:166-209,272-367      MCTreeSearch reset/search/search_node/_expand_node/
 *                                            _play/play_action/_set_node/push_to_queue
 *   games/algos/selfplayworker.py:172-224    SelfPlayer.play_episode and helpers
 *   games/general/modules.py:109-112         network call frame flip (state*player, value*player)
 * The reference has no tests of its own, so this file is pinned against golden
 * vectors produced by running the unmodified reference here (oracle/make_golden.py ->
 * tests/golden/, checked by tests/test_oracle_golden.py).
 *
 * Randomness is injected, not emulated: tie noise / action uniform come from
 * the oracle/spec.py splitmix64 counter stream; Dirichlet noise from a caller table.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -fno-fast-math (see oracle/build.py).
 * All PUCT arithmetic is IEEE double, left to right, no FMA contraction.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OX_MAX_A 9
#define OX_MAX_CELLS 42
#define OX_MAX_MOVES 42

enum { OX_CONNECT4 = 0, OX_TICTACTOE = 1 };
enum { OX_PURPOSE_TIE = 0, OX_PURPOSE_GAMMA = 1, OX_PURPOSE_ACTION = 2, OX_PURPOSE_OPPONENT = 3 };
enum { OX_OPP_MCTS = 0, OX_OPP_LOOKAHEAD = 1, OX_OPP_RANDOM = 2 };

/* ------------------------------------------------------------------ spec stream (oracle/spec.py) */
static uint64_t sm64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    uint64_t z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

uint64_t ox_rng_u64(uint64_t seed, uint64_t game_uid, int tree, int purpose, int ply, uint32_t sim,
                    uint32_t depth, uint64_t idx) {
    uint64_t h = sm64(seed);
    h = sm64(h ^ game_uid);
    h = sm64(h ^ ((uint64_t)(tree & 0xFF) | ((uint64_t)(purpose & 0xFF) << 8) | ((uint64_t)(ply & 0xFFFF) << 16)));
    h = sm64(h ^ ((uint64_t)sim | ((uint64_t)depth << 32)));
    h = sm64(h ^ idx);
    return h;
}

double ox_rng_uniform(uint64_t seed, uint64_t game_uid, int tree, int purpose, int ply, uint32_t sim,
                      uint32_t depth, uint64_t idx) {
    return (double)(ox_rng_u64(seed, game_uid, tree, purpose, ply, sim, depth, idx) >> 11) *
           (1.0 / 9007199254740992.0);
}

/* ------------------------------------------------------------------ envs (array boards, like the reference) */
typedef struct {
    int game, W, H, A;
    int8_t board[OX_MAX_CELLS]; /* [col*H + row]  == numpy board[col,row] */
    int heights[7];
    int episode_over;
} ox_env;

static void env_dims(int game, int* W, int* H, int* A) {
    if (game == OX_CONNECT4) { *W = 7; *H = 6; *A = 7; } else { *W = 3; *H = 3; *A = 9; }
}

void ox_env_reset(ox_env* e, int game) { /* connect4env.py:50-54, tictactoe_env.py:47-50 */
    memset(e, 0, sizeof(*e));
    e->game = game;
    env_dims(game, &e->W, &e->H, &e->A);
}

void ox_env_set_state(ox_env* e, const int8_t* state) { /* connect4env.py:56-58 (episode_over untouched) */
    memcpy(e->board, state, (size_t)(e->W * e->H));
    if (e->game == OX_CONNECT4)
        for (int c = 0; c < e->W; ++c) {
            int h = 0;
            for (int r = 0; r < e->H; ++r) h += abs(e->board[c * e->H + r]);
            e->heights[c] = h;
        }
}

/* functools.reduce(_calc_win_in_a_row, row*player, 0): connect4env.py:85-92 / tictactoe_env.py:76-82 */
static int fold_line(const int* vals, int n, int win) {
    int x = 0;
    for (int i = 0; i < n; ++i) {
        if (x >= win) x = win;
        else if (vals[i] > 0) x = x + vals[i];
        else x = 0;
    }
    return x;
}

/* np.diagonal(M, offset) of a [W,H] matrix: elements M[i, i+offset] */
static int diag_vals(const int8_t* b, int W, int H, int offset, int flip, int player, int* out) {
    int n = 0;
    for (int i = 0; i < W; ++i) {
        int j = i + offset;
        if (j < 0 || j >= H) continue;
        int col = flip ? (W - 1 - i) : i; /* np.flipud reverses axis 0 */
        out[n++] = b[col * H + j] * player;
    }
    return n;
}

static int line_reward(const ox_env* e, int x, int y, int player, int win) {
    /* x = column index (axis 0), y = row index (axis 1) of the last placed cell */
    int vals[8], n, W = e->W, H = e->H;
    n = 0; for (int c = 0; c < W; ++c) vals[n++] = e->board[c * H + y] * player;          /* board[:, y] */
    if (fold_line(vals, n, win) == win) return 1;
    n = 0; for (int r = 0; r < H; ++r) vals[n++] = e->board[x * H + r] * player;          /* board[x, :] */
    if (fold_line(vals, n, win) == win) return 1;
    n = diag_vals(e->board, W, H, y - x, 0, player, vals);                                /* diagonal_1 */
    if (fold_line(vals, n, win) == win) return 1;
    n = diag_vals(e->board, W, H, y - W + x + 1, 1, player, vals);                        /* diagonal_2 */
    if (fold_line(vals, n, win) == win) return 1;
    return 0;
}

/* returns 0 ok, -1 GameOver, -2 ValueError (full column) */
int ox_env_step(ox_env* e, int action, int player, int* reward, int* done) {
    if (e->episode_over) return -1;
    if (e->game == OX_CONNECT4) { /* connect4env.py:29-43 */
        int ph = e->heights[action];
        if (ph < e->H) { e->board[action * e->H + ph] = (int8_t)player; e->heights[action] += 1; }
        else return -2;
        int r = line_reward(e, action, e->heights[action] - 1, player, 4);
        int tot = 0; for (int c = 0; c < e->W; ++c) tot += e->heights[c];
        e->episode_over = (r != 0) || (tot == e->H * e->W);
        *reward = r; *done = e->episode_over;
    } else { /* tictactoe_env.py:23-33: occupied cell is a silent no-op */
        int x = action / e->H, y = action % e->H; /* np.unravel_index(action, (W,H)) */
        if (!e->board[x * e->H + y]) e->board[x * e->H + y] = (int8_t)player;
        int r = line_reward(e, x, y, player, 3);
        int all = 1; for (int i = 0; i < e->W * e->H; ++i) if (!e->board[i]) all = 0;
        e->episode_over = (r != 0) || all;
        *reward = r; *done = e->episode_over;
    }
    return 0;
}

void ox_env_valid_moves(const ox_env* e, uint8_t* valid) { /* connect4env.py:47-48, tictactoe_env.py:42-43 */
    if (e->game == OX_CONNECT4) for (int c = 0; c < e->W; ++c) valid[c] = e->heights[c] < e->H;
    else for (int i = 0; i < 9; ++i) valid[i] = e->board[i] == 0;
}

/* Batch driver for env tests: plays n games of given action lists from the empty board with
 * alternating players (+1 first), recording per-ply outputs.  status: 0 ok,-1 GameOver,-2 ValueError */
void ox_env_playout(int game, int n, int max_plies, const int32_t* actions, const int8_t* first_player,
                    int8_t* boards /*[n][max][cells]*/, int8_t* reward, uint8_t* done,
                    uint8_t* valid /*[n][max][A]*/, int8_t* status) {
    int W, H, A; env_dims(game, &W, &H, &A);
    int cells = W * H;
    for (int g = 0; g < n; ++g) {
        ox_env e; ox_env_reset(&e, game);
        int player = first_player ? first_player[g] : 1;
        for (int t = 0; t < max_plies; ++t) {
            size_t k = (size_t)g * max_plies + t;
            int a = actions[k], r = 0, d = 0;
            int st = (a < 0) ? -3 : ox_env_step(&e, a, player, &r, &d);
            status[k] = (int8_t)st; reward[k] = (int8_t)r; done[k] = (uint8_t)(st == 0 ? d : e.episode_over);
            memcpy(boards + k * cells, e.board, (size_t)cells);
            ox_env_valid_moves(&e, valid + k * A);
            if (st == 0) player = -player;
        }
    }
}

/* ------------------------------------------------------------------ exact integer power (n^k, correctly rounded) */
/* np.power(n, 20.0) (mcts.py:100-101 with temp=1/20) -- computed exactly in 320-bit integer
 * arithmetic and rounded to nearest-even, so CPU and GPU agree and equal a correctly rounded pow. */
double ox_pow_int_exact(uint32_t n, int k) {
    if (k == 0) return 1.0;
    if (n == 0) return 0.0;
    uint32_t limb[40]; int nl = 1; limb[0] = 1;
    for (int i = 0; i < k; ++i) {
        uint64_t carry = 0;
        for (int j = 0; j < nl; ++j) { uint64_t t = (uint64_t)limb[j] * n + carry; limb[j] = (uint32_t)t; carry = t >> 32; }
        if (carry) { if (nl >= 40) return INFINITY; limb[nl++] = (uint32_t)carry; }
    }
    int top = nl - 1; while (top > 0 && limb[top] == 0) --top;
    int hb = 31; while (!((limb[top] >> hb) & 1)) --hb;
    int nbits = top * 32 + hb + 1;
    if (nbits <= 53) { double v = 0; for (int j = top; j >= 0; --j) v = v * 4294967296.0 + (double)limb[j]; return v; }
    int shift = nbits - 53; /* keep top 53 bits */
    uint64_t mant = 0;
    for (int b = nbits - 1; b >= shift; --b) mant = (mant << 1) | ((limb[b / 32] >> (b % 32)) & 1u);
    int half = (limb[(shift - 1) / 32] >> ((shift - 1) % 32)) & 1u;
    int sticky = 0;
    for (int b = shift - 2; b >= 0 && !sticky; --b) sticky |= (limb[b / 32] >> (b % 32)) & 1u;
    if (half && (sticky || (mant & 1))) mant += 1; /* may carry to 2^53: ldexp handles it exactly */
    return ldexp((double)mant, shift);
}

/* ------------------------------------------------------------------ MCNode / MCTreeSearch */
typedef void (*ox_net_fn)(void* user, int tree, const int8_t* state_netframe, float* policy, float* value);

typedef struct {
    int game, sims, evaluate, strong_play;
    int tie_mode;   /* 0: zeros, 1: spec stream                         (mcts.py:355) */
    int noise_mode; /* 0: uniform 1/A, 1: table, 2: generated on the CPU (mcts.py:50)  */
    double alpha;
    uint64_t seed, game_uid;
    const double* noise_table; /* [2][table_moves][A] */
    int table_moves;
    int threads;    /* <= 1: sequential search; K > 1: K cooperative workers with virtual loss (mcts.py:328-331) */
} ox_cfg;

typedef struct {
    int n;              /* mcts.py:28 */
    double w;           /* :29 */
    double p, p_noise;  /* :30,34 */
    int noise_active;   /* :33 */
    int player, valid;  /* :38,39 */
    int virtual_loss;   /* :43 */
    int locked;         /* :47 lock.locked() (threaded search only) */
    int parent;         /* index or -1 */
    int first_child;    /* index of A consecutive children or -1 (anytree children) */
    int has_state;
    double v;
    int8_t state[OX_MAX_CELLS];
} ox_node;

typedef struct {
    int tree, ply;
    int8_t state[OX_MAX_CELLS];
    float probs[OX_MAX_A];
    float q, actual_val;
} ox_record;

typedef struct {
    int tree, ply, action, root_n;
    double root_w;
    int n[OX_MAX_A];
    double w[OX_MAX_A];
} ox_move;

typedef struct {
    ox_cfg cfg;
    int tree_id, W, H, A, cells;
    ox_node* nodes; int n_nodes, cap;
    int root;
    int moves_played;
    ox_net_fn net; void* net_user;
    ox_record temp_memory[OX_MAX_MOVES]; int n_temp;
    long sims_done, net_calls, path_len_sum;
} ox_tree;

static int node_new(ox_tree* t, int parent, int player_if_root, double p, int valid) {
    if (t->n_nodes == t->cap) { t->cap = t->cap ? t->cap * 2 : 1024; t->nodes = (ox_node*)realloc(t->nodes, sizeof(ox_node) * (size_t)t->cap); }
    ox_node* nd = &t->nodes[t->n_nodes];
    memset(nd, 0, sizeof(*nd));
    nd->p = p; nd->valid = valid; nd->parent = parent; nd->first_child = -1;
    nd->player = parent >= 0 ? -1 * t->nodes[parent].player : player_if_root; /* mcts.py:38 */
    return t->n_nodes++;
}

static void create_children(ox_tree* t, int node, const float* probs, const uint8_t* valid) { /* mcts.py:103-107 */
    int first = -1;
    for (int i = 0; i < t->A; ++i) { int c = node_new(t, node, 0, (double)probs[i], valid[i]); if (i == 0) first = c; }
    t->nodes[node].first_child = first;
}

static double node_q(const ox_node* c) { /* mcts.py:59-62 */
    int n_eff = c->n + c->virtual_loss;
    return n_eff ? (c->w - (double)c->virtual_loss) / (double)n_eff : 0.0;
}
static double node_p_eff(const ox_node* c) { /* mcts.py:64-69, x = 0.25 */
    if (c->noise_active) return (c->p_noise * 0.25) + (c->p * (1 - 0.25));
    return c->p;
}
static double node_u(const ox_tree* t, const ox_node* c) { /* mcts.py:71-78, cpuct = 4 */
    const ox_node* par = &t->nodes[c->parent];
    return ((4.0 * node_p_eff(c)) * sqrt((double)(par->n + par->virtual_loss))) / (double)(1 + c->n + c->virtual_loss);
}
static double node_select_prob(const ox_tree* t, const ox_node* c) { /* mcts.py:80-84 */
    return ((double)(-1 * c->player) * node_q(c)) + node_u(t, c);
}
static void node_backup(ox_tree* t, int node, double v) { /* mcts.py:94-98: to the TOP ancestor */
    while (node >= 0) { t->nodes[node].w += v; t->nodes[node].n += 1; node = t->nodes[node].parent; }
}

/* network(s, player): general/modules.py:109-112, inference_proxy.py:21-24 */
static void call_network(ox_tree* t, const int8_t* s, int player, float* probs, double* v) {
    int8_t in[OX_MAX_CELLS];
    for (int i = 0; i < t->cells; ++i) in[i] = (int8_t)(s[i] * player);
    float value = 0.f;
    t->net(t->net_user, t->tree_id, in, probs, &value);
    t->net_calls++;
    *v = (double)value * (double)player;
}

/* mcts.py:301-321 */
static int expand_node(ox_tree* t, int parent, int action, int player, double* v_out) {
    ox_env env; ox_env_reset(&env, t->cfg.game);
    ox_env_set_state(&env, t->nodes[parent].state);
    int r = 0, done = 0;
    ox_env_step(&env, action, player, &r, &done);
    r = r * player;
    int child = t->nodes[parent].first_child + action;
    double v;
    if (done) {
        if (t->cfg.strong_play) {
            int num_steps = 1; for (int i = 0; i < t->cells; ++i) num_steps += abs(t->nodes[parent].state[i]);
            v = (1.18 - ((double)(9 * num_steps) / 350.0)) * (double)r;
        } else v = (double)r;
    } else {
        float probs[OX_MAX_A]; uint8_t valid[OX_MAX_A];
        call_network(t, env.board, t->nodes[parent].player, probs, &v);
        ox_env_valid_moves(&env, valid);
        create_children(t, child, probs, valid);
    }
    memcpy(t->nodes[child].state, env.board, (size_t)t->cells);
    t->nodes[child].has_state = 1;
    *v_out = v;
    return child;
}

void ox_tree_reset(ox_tree* t, int player) { /* mcts.py:166-174 */
    t->n_nodes = 0;
    ox_env env; ox_env_reset(&env, t->cfg.game);
    float probs[OX_MAX_A]; double v; uint8_t valid[OX_MAX_A];
    call_network(t, env.board, 1, probs, &v);
    int root = node_new(t, -1, player, 0.0, 1);
    t->nodes[root].v = v; t->nodes[root].has_state = 1;
    memcpy(t->nodes[root].state, env.board, (size_t)t->cells);
    ox_env_valid_moves(&env, valid);
    create_children(t, root, probs, valid);
    t->root = root; t->moves_played = 0; t->n_temp = 0;
}

ox_tree* ox_tree_new(const ox_cfg* cfg, int tree_id, ox_net_fn net, void* user) {
    ox_tree* t = (ox_tree*)calloc(1, sizeof(ox_tree));
    t->cfg = *cfg; t->tree_id = tree_id; t->net = net; t->net_user = user;
    env_dims(cfg->game, &t->W, &t->H, &t->A); t->cells = t->W * t->H;
    return t;
}
void ox_tree_free(ox_tree* t) { if (t) { free(t->nodes); free(t); } }

static int root_ply(const ox_tree* t) { int s = 0; for (int i = 0; i < t->cells; ++i) s += abs(t->nodes[t->root].state[i]); return s; }

/* ---- Dirichlet noise generated on the CPU (noise_mode 2; same draw schedule as csrc/spx_rng.cuh,
 * libm transcendental functions so NOT bit-comparable with the GPU: parity tests inject tables). */
static double gamma_variate(const ox_cfg* c, int tree, int ply, int action, double alpha) {
    uint32_t attempt = 0;
#define U(j) ox_rng_uniform(c->seed, c->game_uid, tree, OX_PURPOSE_GAMMA, ply, attempt, 0, (uint64_t)action * 4 + (j))
    if (alpha == 1.0) return -log(1.0 - U(0));
    if (alpha < 1.0) {
        for (;; ++attempt) {
            double Uv = U(0), V = -log(1.0 - U(1));
            if (Uv <= 1.0 - alpha) { double X = pow(Uv, 1.0 / alpha); if (X <= V) return X; }
            else { double Y = -log((1.0 - Uv) / alpha); double X = pow(1.0 - alpha + alpha * Y, 1.0 / alpha); if (X <= V + Y) return X; }
        }
    }
    double b = alpha - 1.0 / 3.0, cc = 1.0 / sqrt(9.0 * b);
    for (;; ++attempt) {
        double X = sqrt(-2.0 * log(1.0 - U(0))) * cos(6.283185307179586 * U(1));
        double V = 1.0 + cc * X;
        if (V <= 0.0) continue;
        V = V * V * V;
        double Uv = U(2);
        if (Uv < 1.0 - 0.0331 * (X * X) * (X * X)) return b * V;
        if (log(Uv) < 0.5 * X * X + b * (1.0 - V + log(V))) return b * V;
    }
#undef U
}

static void add_noise(ox_tree* t) { /* mcts.py:49-53 */
    ox_node* root = &t->nodes[t->root];
    double d[OX_MAX_A];
    if (t->cfg.noise_mode == 1 && t->cfg.noise_table) {
        const double* row = t->cfg.noise_table + ((size_t)t->tree_id * t->cfg.table_moves + t->moves_played) * t->A;
        for (int i = 0; i < t->A; ++i) d[i] = row[i];
    } else if (t->cfg.noise_mode == 2) {
        double acc = 0; int ply = root_ply(t);
        for (int i = 0; i < t->A; ++i) { d[i] = gamma_variate(&t->cfg, t->tree_id, ply, i, t->cfg.alpha); acc += d[i]; }
        double inv = 1.0 / acc;
        for (int i = 0; i < t->A; ++i) d[i] = d[i] * inv;
    } else for (int i = 0; i < t->A; ++i) d[i] = 1.0 / (double)t->A;
    for (int i = 0; i < t->A; ++i) { ox_node* c = &t->nodes[root->first_child + i]; c->noise_active = 1; c->p_noise = d[i]; }
}
static void remove_noise(ox_tree* t) { /* mcts.py:55-57 */
    ox_node* root = &t->nodes[t->root];
    for (int i = 0; i < t->A; ++i) t->nodes[root->first_child + i].noise_active = 0;
}

/* live counters for the time-boxed CPU baseline (bench.py): [0] sims, [1] moves; may point into shared memory */
static volatile long* g_live = 0;
void ox_set_live_counters(long* p) { g_live = p; }

/* mcts.py:340-367 (sequential mode: thread_count == 1) */
static void search_node(ox_tree* t, int ply, uint32_t sim) {
    int node = t->root;
    int node_list[OX_MAX_MOVES + 2], depth = 0;
    for (;;) {
        node_list[depth] = node;
        t->nodes[node].virtual_loss += 1;
        double scores[OX_MAX_A]; int all_bad = 1;
        for (int a = 0; a < t->A; ++a) {
            const ox_node* c = &t->nodes[t->nodes[node].first_child + a];
            scores[a] = c->valid ? node_select_prob(t, c) : -10000000000.0;
            if (!(scores[a] < -100000)) all_bad = 0;
        }
        if (all_bad) return; /* mcts.py:349-354: virtual loss deliberately NOT removed */
        int best = 0; double best_s = 0;
        for (int a = 0; a < t->A; ++a) { /* np.argmax(select_probs + 1e-6*rand(A)): first maximum wins */
            double noise = t->cfg.tie_mode ? ox_rng_uniform(t->cfg.seed, t->cfg.game_uid, t->tree_id, OX_PURPOSE_TIE, ply, sim, (uint32_t)depth, (uint64_t)a) : 0.0;
            double s = scores[a] + (0.000001 * noise);
            if (a == 0 || s > best_s) { best = a; best_s = s; }
        }
        int child = t->nodes[node].first_child + best;
        depth++;
        if (t->nodes[child].first_child < 0) { /* is_leaf: unexpanded OR terminal */
            double v;
            int nd = expand_node(t, node, best, t->nodes[node].player, &v);
            node_backup(t, nd, v);
            t->nodes[nd].v = v;
            for (int i = 0; i < depth; ++i) t->nodes[node_list[i]].virtual_loss -= 1;
            break;
        } else node = child;
    }
    t->path_len_sum += depth;
}

/* ---- threaded search (mcts.py:328-331: ThreadPoolExecutor(thread_count) runs `iterations` search_node tasks) restated as ONE
 * legal interleaving of those threads, the cooperative round-robin schedule: worker 0 runs search_node tasks until it blocks in
 * a network call (a task that needs none -- terminal leaf, or the "all states in use" return -- completes and the worker takes
 * the next task), then worker 1, ..., worker K-1; the K pending evaluations come back together; worker 0 resumes (create_children,
 * backup, virtual loss removed, lock released) and runs on, then worker 1, ...  until all tasks are done.  While a worker waits
 * it holds the lock of the child it expands (mcts.py:358,362), so `valid` is false for it (:86-88), and its virtual losses stay
 * on the path (:345).  tests/golden/threaded.json holds the unmodified reference run under exactly this schedule. */
typedef struct {
    int waiting;                     /* blocked in the network call of `child` */
    int node_list[OX_MAX_MOVES + 2], depth;
    int parent, action, child;
    float probs[OX_MAX_A]; double v; uint8_t valid[OX_MAX_A];
    int8_t child_state[OX_MAX_CELLS];
} ox_worker;

/* one search_node task of worker w up to its network call; returns 1 if it blocks there, 0 if the task completed */
static int worker_run_task(ox_tree* t, ox_worker* w, int ply, uint32_t sim) {
    int node = t->root;
    w->depth = 0;
    for (;;) {
        w->node_list[w->depth] = node;
        t->nodes[node].virtual_loss += 1;
        double scores[OX_MAX_A]; int all_bad = 1;
        for (int a = 0; a < t->A; ++a) {
            const ox_node* c = &t->nodes[t->nodes[node].first_child + a];
            scores[a] = (c->valid && !c->locked) ? node_select_prob(t, c) : -10000000000.0;   /* valid: mcts.py:86-88 */
            if (!(scores[a] < -100000)) all_bad = 0;
        }
        if (all_bad) return 0; /* mcts.py:349-354: the task ends, virtual loss deliberately NOT removed */
        int best = 0; double best_s = 0;
        for (int a = 0; a < t->A; ++a) {
            double noise = t->cfg.tie_mode ? ox_rng_uniform(t->cfg.seed, t->cfg.game_uid, t->tree_id, OX_PURPOSE_TIE, ply, sim, (uint32_t)w->depth, (uint64_t)a) : 0.0;
            double s = scores[a] + (0.000001 * noise);
            if (a == 0 || s > best_s) { best = a; best_s = s; }
        }
        int child = t->nodes[node].first_child + best;
        w->depth++;
        if (t->nodes[child].first_child < 0) { /* is_leaf */
            t->nodes[child].locked = 1;                                   /* :358 */
            ox_env env; ox_env_reset(&env, t->cfg.game);                 /* _expand_node :301-321 up to the network call */
            ox_env_set_state(&env, t->nodes[node].state);
            int r = 0, done = 0, player = t->nodes[node].player;
            ox_env_step(&env, best, player, &r, &done);
            r = r * player;
            t->path_len_sum += w->depth;
            if (done) {
                double v;
                if (t->cfg.strong_play) {
                    int num_steps = 1; for (int i = 0; i < t->cells; ++i) num_steps += abs(t->nodes[node].state[i]);
                    v = (1.18 - ((double)(9 * num_steps) / 350.0)) * (double)r;
                } else v = (double)r;
                memcpy(t->nodes[child].state, env.board, (size_t)t->cells); t->nodes[child].has_state = 1;
                node_backup(t, child, v); t->nodes[child].v = v;
                t->nodes[child].locked = 0;
                for (int i = 0; i < w->depth; ++i) t->nodes[w->node_list[i]].virtual_loss -= 1;
                return 0;
            }
            call_network(t, env.board, player, w->probs, &w->v);          /* the answer is consumed when the worker resumes */
            ox_env_valid_moves(&env, w->valid);
            memcpy(w->child_state, env.board, (size_t)t->cells);
            w->parent = node; w->action = best; w->child = child; w->waiting = 1;
            return 1;
        }
        node = child;
    }
}

static void worker_resume(ox_tree* t, ox_worker* w) {   /* _expand_node after the network call, backup, unlock, virtual loss removed */
    create_children(t, w->child, w->probs, w->valid);
    memcpy(t->nodes[w->child].state, w->child_state, (size_t)t->cells); t->nodes[w->child].has_state = 1;
    node_backup(t, w->child, w->v); t->nodes[w->child].v = w->v;
    t->nodes[w->child].locked = 0;
    for (int i = 0; i < w->depth; ++i) t->nodes[w->node_list[i]].virtual_loss -= 1;
    w->waiting = 0;
}

static void search_threaded(ox_tree* t, int ply) {
    const int K = t->cfg.threads;
    ox_worker* ws = (ox_worker*)calloc((size_t)K, sizeof(ox_worker));
    int next_task = 0, any = 1;
    while (any) {
        any = 0;
        for (int k = 0; k < K; ++k) {
            ox_worker* w = &ws[k];
            if (w->waiting) { worker_resume(t, w); t->sims_done++; if (g_live) g_live[0]++; }
            while (!w->waiting && next_task < t->cfg.sims) {
                const uint32_t sim = (uint32_t)next_task++;
                if (!worker_run_task(t, w, ply, sim)) { t->sims_done++; if (g_live) g_live[0]++; }
            }
            if (w->waiting) any = 1;
        }
    }
    free(ws);
}

void ox_tree_search(ox_tree* t) { /* mcts.py:323-338 */
    add_noise(t);
    int ply = root_ply(t);
    if (t->cfg.threads > 1) search_threaded(t, ply);
    else for (int i = 0; i < t->cfg.sims; ++i) { search_node(t, ply, (uint32_t)i); t->sims_done++; if (g_live) g_live[0]++; }
    remove_noise(t);
}

/* mcts.py:272-299.  Returns the action; appends a record unless the ValueError fallback fires. */
int ox_tree_play(ox_tree* t, ox_move* log) {
    ox_node* root = &t->nodes[t->root];
    double temp = 1.0;                       /* mcts.py:182-183: always 1 */
    if (t->cfg.evaluate) temp = temp / 20;   /* :273-274 */
    double inv_t = 1 / temp;
    double pp[OX_MAX_A], sum = 0;
    for (int a = 0; a < t->A; ++a) {
        int n = t->nodes[root->first_child + a].n;
        if (inv_t == 1.0) pp[a] = (double)n;
        else if (inv_t == floor(inv_t) && inv_t <= 64) pp[a] = ox_pow_int_exact((uint32_t)n, (int)inv_t);
        else pp[a] = pow((double)n, inv_t);
        sum = sum + pp[a];
    }
    int ply = root_ply(t);
    if (log) {
        log->tree = t->tree_id; log->ply = ply; log->root_n = root->n; log->root_w = root->w;
        for (int a = 0; a < t->A; ++a) { log->n[a] = t->nodes[root->first_child + a].n; log->w[a] = t->nodes[root->first_child + a].w; }
    }
    int action;
    if (sum == 0 || sum != sum || isinf(sum)) { /* p has NaN -> ValueError -> argmax of n, no record (:290-295) */
        action = 0;
        for (int a = 1; a < t->A; ++a) if (t->nodes[root->first_child + a].n > t->nodes[root->first_child + action].n) action = a;
    } else {
        double probs[OX_MAX_A], cdf[OX_MAX_A], acc = 0;
        for (int a = 0; a < t->A; ++a) { probs[a] = pp[a] / sum; acc = acc + probs[a]; cdf[a] = acc; }
        for (int a = 0; a < t->A; ++a) cdf[a] = cdf[a] / cdf[t->A - 1]; /* numpy choice: cdf /= cdf[-1] */
        double u = ox_rng_uniform(t->cfg.seed, t->cfg.game_uid, t->tree_id, OX_PURPOSE_ACTION, ply, 0, 0, 0);
        action = 0; while (action < t->A && cdf[action] <= u) ++action; /* searchsorted(side='right') */
        if (action >= t->A) action = t->A - 1;
        ox_record* rec = &t->temp_memory[t->n_temp++];
        memset(rec, 0, sizeof(*rec));
        rec->tree = t->tree_id; rec->ply = ply;
        memcpy(rec->state, root->state, (size_t)t->cells);
        for (int a = 0; a < t->A; ++a) rec->probs[a] = (float)probs[a];
        rec->q = (float)node_q(root);
    }
    if (log) log->action = action;
    if (g_live) g_live[1]++;
    t->moves_played += 1;
    return action;
}

void ox_tree_play_action(ox_tree* t, int action) { /* mcts.py:188-209 */
    int node = t->nodes[t->root].first_child + action;
    if (t->nodes[node].n == 0) {
        double v;
        node = expand_node(t, t->root, action, t->nodes[t->root].player, &v);
        node_backup(t, node, v);
        t->nodes[node].v = v;
    }
    t->root = node;
}

void ox_tree_root_stats(const ox_tree* t, int32_t* n, double* w, uint8_t* valid, int32_t* root_n, double* root_w,
                        double* q, int32_t* player) {
    const ox_node* root = &t->nodes[t->root];
    for (int a = 0; a < t->A; ++a) {
        const ox_node* c = &t->nodes[root->first_child + a];
        n[a] = c->n; w[a] = c->w; valid[a] = (uint8_t)c->valid;
    }
    *root_n = root->n; *root_w = root->w; *q = node_q(root); *player = root->player;
}
long ox_tree_counter(const ox_tree* t, int which) {
    return which == 0 ? t->sims_done : which == 1 ? t->net_calls : which == 2 ? t->path_len_sum : t->n_nodes;
}

/* ------------------------------------------------------------------ SelfPlayer.play_episode */
typedef struct {
    int reward, n_moves, n_records;
    long sims, net_calls, path_len_sum;
    int8_t final_state[OX_MAX_CELLS];
    ox_move moves[OX_MAX_MOVES + 1];
    ox_record records[OX_MAX_MOVES + 1];
} ox_episode;

typedef struct { ox_tree* t[2]; ox_env env; ox_episode* out; } ox_player;

/* selfplayworker.py:221-224 */
static void play_move(ox_player* sp, int a, int player, int* r, int* done) {
    ox_tree_play_action(sp->t[0], a);
    ox_tree_play_action(sp->t[1], a);
    ox_env_step(&sp->env, a, player, r, done);
}
/* selfplayworker.py:207-219 */
static void get_and_play_moves(ox_player* sp, int player, int* r, int* done) {
    ox_tree* t = player == 1 ? sp->t[0] : sp->t[1];
    ox_tree_search(t);
    int a = ox_tree_play(t, &sp->out->moves[sp->out->n_moves]);
    sp->out->n_moves++;
    play_move(sp, a, player, r, done);
    if (player != 1) *r = *r * player;
}

/* selfplayworker.py:172-194 with update=True; each tree may use its own net (evaluate mode) */
int ox_play_episode(const ox_cfg* cfg, int swap_sides, ox_net_fn net0, void* user0, ox_net_fn net1, void* user1,
                    ox_episode* out) {
    memset(out, 0, sizeof(*out));
    ox_player sp; sp.out = out;
    sp.t[0] = ox_tree_new(cfg, 0, net0, user0);
    sp.t[1] = ox_tree_new(cfg, 1, net1, user1);
    ox_env_reset(&sp.env, cfg->game);
    ox_tree_reset(sp.t[0], swap_sides ? -1 : 1);
    ox_tree_reset(sp.t[1], swap_sides ? 1 : -1);
    int r = 0, done = 0;
    int max_moves = sp.env.W * sp.env.H;
    if (swap_sides) get_and_play_moves(&sp, -1, &r, &done);
    for (int i = 0; i < max_moves; ++i) { /* play_round: :196-204 */
        get_and_play_moves(&sp, 1, &r, &done);
        if (!done) get_and_play_moves(&sp, -1, &r, &done);
        if (done) break;
    }
    out->reward = r;
    memcpy(out->final_state, sp.env.board, sizeof(out->final_state));
    /* push_to_queue (mcts.py:225-232): policy gets r, opposing gets -r; policy's records first */
    for (int k = 0; k < 2; ++k) {
        ox_tree* t = sp.t[k];
        for (int i = 0; i < t->n_temp; ++i) {
            ox_record rec = t->temp_memory[i];
            rec.actual_val = (float)(k == 0 ? r : r * -1);
            out->records[out->n_records++] = rec;
        }
        out->sims += t->sims_done; out->net_calls += t->net_calls; out->path_len_sum += t->path_len_sum;
    }
    ox_tree_free(sp.t[0]); ox_tree_free(sp.t[1]);
    return 0;
}

/* ------------------------------------------------------------------ hard-coded opponents (general/hardcoded_players.py)
 * The opponent keeps its OWN env in its OWN frame: play_action(a, player) steps it with the player the SelfPlayer
 * passes (selfplayworker.py:221-224: the opponent's moves arrive as +1, the policy's as -1). */
typedef struct { int kind, player; ox_env env; } ox_hardcoded;

static int hardcoded_move(ox_hardcoded* h, const ox_cfg* cfg) {
    uint8_t valid[OX_MAX_A]; int moves[OX_MAX_A], n = 0;
    ox_env_valid_moves(&h->env, valid);
    for (int a = 0; a < h->env.A; ++a) if (valid[a]) moves[n++] = a;      /* possible_moves, :18,47 */
    if (h->kind == OX_OPP_LOOKAHEAD) {                                      /* OneStepLookahead.__call__ :15-30 */
        for (int pass = 0; pass < 2; ++pass) {
            int who = pass == 0 ? h->player : -h->player;                   /* "can win" uses self.player, then the block */
            for (int i = 0; i < n; ++i) {
                ox_env t = h->env; t.episode_over = 0;                      /* test_env.set_state(copy(state)) on a fresh env */
                int r = 0, done = 0;
                ox_env_step(&t, moves[i], who, &r, &done);
                if (done) return moves[i];
            }
        }
    }
    int ply = 0; for (int i = 0; i < h->env.W * h->env.H; ++i) ply += abs(h->env.board[i]);
    double u = ox_rng_uniform(cfg->seed, cfg->game_uid, 1, OX_PURPOSE_OPPONENT, ply, 0, 0, 0);
    int idx = (int)(u * (double)n);                                         /* random.choice, injected */
    if (idx >= n) idx = n - 1;
    return moves[idx];
}

/* SelfPlayer.play_episode with update=False, policy = MCTreeSearch (tree 0), opposing = hard-coded player */
int ox_play_episode_vs(const ox_cfg* cfg, int swap_sides, int kind, ox_net_fn net0, void* user0, ox_episode* out) {
    memset(out, 0, sizeof(*out));
    ox_tree* t0 = ox_tree_new(cfg, 0, net0, user0);
    ox_hardcoded h; h.kind = kind;
    ox_env env; ox_env_reset(&env, cfg->game);
    ox_tree_reset(t0, swap_sides ? -1 : 1);
    h.player = swap_sides ? 1 : -1;                                         /* reset(player=...) :32-34 (Random ignores it) */
    ox_env_reset(&h.env, cfg->game);
    int r = 0, done = 0, max_moves = env.W * env.H;
    for (int turn = swap_sides ? -1 : 0; turn < 2 * max_moves && !done; ++turn) {
        int player = (turn < 0 || (turn & 1)) ? -1 : 1;                     /* swap: opponent first, then policy/opponent rounds */
        int a;
        if (player == 1) {
            ox_tree_search(t0);
            a = ox_tree_play(t0, &out->moves[out->n_moves]);
        } else {
            a = hardcoded_move(&h, cfg);
            ox_move* m = &out->moves[out->n_moves];
            memset(m, 0, sizeof(*m)); m->tree = 1; m->action = a;
            for (int i = 0; i < env.W * env.H; ++i) m->ply += abs(env.board[i]);
        }
        out->n_moves++;
        ox_tree_play_action(t0, a);                                         /* play_move :221-224 */
        { int rr, dd; ox_env_step(&h.env, a, player * -1, &rr, &dd); }
        ox_env_step(&env, a, player, &r, &done);
        if (player != 1) r = r * player;
    }
    out->reward = r;
    memcpy(out->final_state, env.board, sizeof(out->final_state));
    out->sims = t0->sims_done; out->net_calls = t0->net_calls; out->path_len_sum = t0->path_len_sum;
    ox_tree_free(t0);
    return 0;
}

/* ------------------------------------------------------------------ built-in networks for tests */
typedef struct { int game; uint64_t net_seed; long calls; } ox_hashnet_state;

static void board_bits(const int8_t* s, int game, uint64_t* own, uint64_t* opp) { /* oracle/spec.py board_to_bits */
    int W, H, A; env_dims(game, &W, &H, &A);
    int stride = game == OX_CONNECT4 ? 7 : 3;
    *own = *opp = 0;
    for (int c = 0; c < W; ++c) for (int r = 0; r < H; ++r) {
        int v = s[c * H + r];
        if (v == 1) *own |= 1ULL << (c * stride + r); else if (v == -1) *opp |= 1ULL << (c * stride + r);
    }
}

void ox_hashnet_bits(uint64_t own, uint64_t opp, int A, uint64_t net_seed, float* policy, float* value) { /* spec.hashnet */
    uint64_t k = sm64(own ^ sm64(opp + net_seed));
    uint32_t r[OX_MAX_A]; uint32_t tot = 0;
    for (int i = 0; i < A; ++i) { r[i] = (uint32_t)(sm64(k ^ (uint64_t)(i + 1)) >> 48) + 1; tot += r[i]; }
    float ftot = (float)tot;
    for (int i = 0; i < A; ++i) policy[i] = (float)r[i] / ftot;
    float vv = (float)(uint32_t)(sm64(k ^ 0xFFULL) >> 48) - 32768.0f;
    *value = vv / 81920.0f;
}

void ox_hashnet(void* user, int tree, const int8_t* state, float* policy, float* value) {
    (void)tree;
    ox_hashnet_state* hs = (ox_hashnet_state*)user;
    uint64_t own, opp; board_bits(state, hs->game, &own, &opp);
    int W, H, A; env_dims(hs->game, &W, &H, &A);
    ox_hashnet_bits(own, opp, A, hs->net_seed, policy, value);
    hs->calls++;
}

/* Replays logged network outputs in call order per tree and checks that the oracle asks for the
 * same leaf states the engine evaluated ("given identical network outputs"). */
typedef struct {
    int game, A;
    long n[2], cursor[2], mismatches, overruns;
    const uint64_t* own[2]; const uint64_t* opp[2];
    const float* policy[2]; const float* value[2];
} ox_replay_state;

void ox_replaynet(void* user, int tree, const int8_t* state, float* policy, float* value) {
    ox_replay_state* rs = (ox_replay_state*)user;
    long i = rs->cursor[tree]++;
    if (i >= rs->n[tree]) { rs->overruns++; for (int a = 0; a < rs->A; ++a) policy[a] = 1.0f / (float)rs->A; *value = 0.f; return; }
    uint64_t own, opp; board_bits(state, rs->game, &own, &opp);
    if (own != rs->own[tree][i] || opp != rs->opp[tree][i]) rs->mismatches++;
    for (int a = 0; a < rs->A; ++a) policy[a] = rs->policy[tree][i * rs->A + a];
    *value = rs->value[tree][i];
}

size_t ox_sizeof(int what) {
    return what == 0 ? sizeof(ox_cfg) : what == 1 ? sizeof(ox_episode) : what == 2 ? sizeof(ox_record) : what == 3 ? sizeof(ox_move)
         : what == 4 ? sizeof(ox_hashnet_state) : what == 5 ? sizeof(ox_replay_state) : sizeof(ox_env);
}

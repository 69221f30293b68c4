// spx_engine.cu -- batched AlphaZero self-play search engine for sm_100a.
//
// One warp per game slot.  Every slot is a small state machine that replays, for one game,
// SelfPlayer.play_episode (selfplayworker.py:172-224) over two MCTreeSearch trees
// (mcts.py:116-367): reset -> search (S x search_node) -> _play -> play_action on both trees ->
// env.step -> ... -> push_to_queue.  Whenever the reference would call the network
// (mcts.py:168,316) the slot parks, writes the leaf position into the dense leaf batch and resumes on
// the next spx_advance() with the network outputs.  Sims that end on a terminal child
// (mcts.py:357-365 with `done`) need no network and are resolved inside the same launch.
//
// Tree storage: a flat per-tree bump-allocated node pool in HBM.  A node holds the statistics of its
// A child edges (MCNode.n/.w/.p of the children, mcts.py:28-30) contiguously:
//     w f64[A] | n i32[A] | p f32[A] | child i32[A] | meta u32 | own u64 | opp u64      (160 B, A=7)
// so one PUCT level is a single 160-byte coalesced read by lanes 0..A-1, a warp-shuffle argmax picks
// the child (first maximum wins, like np.argmax), and backup touches one (n,w) pair per level with
// every level handled by a different lane.  All PUCT arithmetic is IEEE fp64 with explicit
// round-to-nearest intrinsics (no FMA contraction) to reproduce mcts.py:59-84 bit for bit.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <atomic>
#include <new>

#include <cub/device/device_radix_sort.cuh>

#include "spx_advance.cuh"

namespace spx {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

int set_err(int code, const char* fmt, const char* detail) {
    snprintf(g_err, sizeof(g_err), fmt, detail ? detail : "");
    return code;
}
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

#define SPX_CUDA(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) return spx::set_err(SPX_E_CUDA, #expr ": %s", cudaGetErrorString(_e)); \
    } while (0)

// ------------------------------------------------------------------------------------------------ the tick kernel
#ifndef SPX_ADV_MINB
#define SPX_ADV_MINB 4   // 128 registers: advance_game needs 96 (Connect4) / 110 (TicTacToe) and must not spill on its select chain
#endif
template <int GAME>
__global__ void __launch_bounds__(128, SPX_ADV_MINB) advance_kernel(EngineDev E, const float* __restrict__ policy_in,
                                                      const float* __restrict__ value_in) {
    const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (g >= E.cfg.n_games) return;
    const int lane = threadIdx.x & 31;
    const AdvPre pre = advance_prefetch<GAME>(E, g, lane);
    // the outputs of the evaluation this slot asked for (null before the first evaluation: nothing is pending then)
    const float my_p = (policy_in && lane < Rules<GAME>::A) ? policy_in[(size_t)g * Rules<GAME>::A + lane] : 0.f;
    const float v = value_in ? value_in[g] : 0.f;
    u64 own, opp;
    advance_game<GAME>(E, g, lane, pre, my_p, v, E.cfg.max_sims_per_tick, false, g == 0, own, opp);
}

// the same tick with the evaluation cache: E.ecache / E.ecache_tag are set by spx_advance for this launch
template <int GAME>
__global__ void __launch_bounds__(128, 3) advance_kernel_cached(EngineDev E, const float* __restrict__ policy_in, const float* __restrict__ value_in) {
    const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (g >= E.cfg.n_games) return;
    const int lane = threadIdx.x & 31;
    const AdvPre pre = advance_prefetch<GAME>(E, g, lane);
    const float my_p = (policy_in && lane < Rules<GAME>::A) ? policy_in[(size_t)g * Rules<GAME>::A + lane] : 0.f;
    const float v = value_in ? value_in[g] : 0.f;
    u64 own, opp;
    advance_game<GAME, false, true>(E, g, lane, pre, my_p, v, E.cfg.max_sims_per_tick, false, g == 0, own, opp);
}

// the same tick for an engine created with search_threads = K > 1 (leaf slot of worker k of game g = g * K + k)
template <int GAME>
__global__ void __launch_bounds__(128, 2) advance_kernel_threaded(EngineDev E, const float* __restrict__ policy_in, const float* __restrict__ value_in) {
    const int g = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (g >= E.cfg.n_games) return;
    const int lane = threadIdx.x & 31;
    const AdvPre pre = advance_prefetch<GAME>(E, g, lane);
    const size_t ls = (size_t)g * E.K;
    const float my_p = (policy_in && lane < Rules<GAME>::A) ? policy_in[ls * Rules<GAME>::A + lane] : 0.f;
    const float v = value_in ? value_in[ls] : 0.f;
    u64 own, opp;
    advance_game<GAME, true>(E, g, lane, pre, my_p, v, E.cfg.max_sims_per_tick, false, g == 0, own, opp, policy_in, value_in);
}

__global__ void sqrt_table_kernel(double* t) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < sf::SQRT_TABLE) t[i] = __dsqrt_rn((double)i);
}

// spx_softf64_selftest: the integer-pipe arithmetic of spx_softf64.cuh against the FP64 instructions, op by op, on operand pairs drawn
// from a counter stream (exponents kept close so that sums cancel and round, mantissas full random; every 16th pair is a special:
// zeros, equal magnitudes, powers of two, the largest / smallest mantissas)
__global__ void softf64_selftest_kernel(unsigned long long n, unsigned long long seed, unsigned long long* mism, const double* table) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned long long bad[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const u64 h0 = sm64(seed ^ (i * 4 + 0)), h1 = sm64(seed ^ (i * 4 + 1)), h2 = sm64(seed ^ (i * 4 + 2)), h3 = sm64(seed ^ (i * 4 + 3));
        const int ea = 1023 - 40 + (int)(h2 % 80), eb = ea - 70 + (int)((h2 >> 8) % 140);
        u64 xa = (h0 & 0x800FFFFFFFFFFFFFULL) | ((u64)ea << 52), xb = (h1 & 0x800FFFFFFFFFFFFFULL) | ((u64)eb << 52);
        if ((i & 15) == 0) {
            const int kind = (int)((i >> 4) % 7);
            if (kind == 0) xa &= sf::SIGN;                                   // +-0 and something
            if (kind == 1) { xa &= sf::SIGN; xb &= sf::SIGN; }               // +-0 and +-0
            if (kind == 2) xb = xa ^ sf::SIGN;                               // x + (-x)
            if (kind == 3) xa &= ~sf::MASK52;                                // a power of two
            if (kind == 4) xa |= sf::MASK52;                                 // all-ones mantissa
            if (kind == 5) { xb = (xb & ~(0x7FFULL << 52)) | ((u64)(ea - 1 + (int)(h3 & 3)) << 52); xb ^= ((xa ^ xb) & sf::SIGN) ^ sf::SIGN; }   // near-cancellation
            if (kind == 6) { xa = (xa & ~sf::MASK52) | (h0 & 0x7); xb = (xb & ~sf::MASK52) | (h1 & 0x3); }                                       // short mantissas: exact ties
        }
        const double a = sf::dbl(xa), b = sf::dbl(xb);
        if (sf::bits(sf::mul(a, b)) != sf::bits(__dmul_rn(a, b))) bad[0]++;
        if (sf::bits(sf::add(a, b)) != sf::bits(__dadd_rn(a, b))) bad[1]++;
        const unsigned k = (h3 & 1) ? (unsigned)(1 + (h3 >> 8) % 70000) : (unsigned)(1 + ((h3 >> 8) & 0x7FFFFFFE) % 0x7FFFFFFEu);
        if (sf::bits(sf::div_u32(a, k)) != sf::bits(__ddiv_rn(a, (double)k))) bad[2]++;
        const unsigned m = (unsigned)(h3 >> 40) % (2 * sf::SQRT_TABLE);
        if (sf::bits(sf::sqrt_u32(m, table)) != sf::bits(__dsqrt_rn((double)m))) bad[3]++;
        const float f = __uint_as_float((unsigned)h3 & ((i & 31) == 1 ? 0x807FFFFFu : 0xFFFFFFFFu));
        if (f == f && sf::bits(sf::from_f32(f)) != sf::bits((double)f)) bad[4]++;
        const u64 k53 = (h2 >> 11) >> (h3 % 53);
        if (sf::bits(sf::from_u53(k53)) != sf::bits((double)k53 * (1.0 / 9007199254740992.0))) bad[5]++;
        if (sf::gt(a, b) != (a > b) || sf::eq(a, b) != (a == b) || sf::gt(b, a) != (b > a)) bad[6]++;
        if (sf::bits(sf::mul_pow2(a, 2)) != sf::bits(__dmul_rn(a, 4.0)) || sf::bits(sf::mul_pow2(a, -2)) != sf::bits(__dmul_rn(a, 0.25)) ||
            sf::bits(sf::neg_if(a, true)) != sf::bits(__dmul_rn(a, -1.0))) bad[7]++;
    }
    for (int j = 0; j < 8; ++j) if (bad[j]) atomicAdd(mism + j, bad[j]);
}

__global__ void reset_kernel(EngineDev E) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= E.cfg.n_games) return;
    GameState s;
    memset(&s, 0, sizeof(s));
    s.game_index = (u64)(E.cfg.slot_offset + g);
    s.swap = (int)(s.game_index & 1ULL);
    s.phase = ((long long)s.game_index < E.cfg.games_target) ? PH_RESET : PH_IDLE;
    s.tree[0].root = s.tree[1].root = -1;
    E.games[g] = s;
    for (int k = 0; k < E.K; ++k) {
        const size_t l = (size_t)g * E.K + k;
        E.leaf_own[l] = 0; E.leaf_opp[l] = 0; E.needs_eval[l] = 0; E.net_id[l] = 0;
        if (E.workers) memset(&E.workers[l], 0, sizeof(Worker));
    }
    E.ext_action[g] = -1; E.own_action[2 * g] = 0; E.own_action[2 * g + 1] = -1;
    if (g == 0) { *E.rec_count = 0; *E.res_count = 0; *E.rec_dropped = 0; *E.ticks = 0; }
}

template <int GAME>
__global__ void root_stats_kernel(EngineDev E, int tree, int* n, double* w, int* root_n, double* root_w, unsigned short* valid) {
    typedef NodeLayout<GAME> L;
    constexpr int A = Rules<GAME>::A;
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= E.cfg.n_games) return;
    const TreeState ts = E.games[g].tree[tree];
    const char* nd = ts.root >= 0 ? E.pool + (((size_t)g * 2 + tree) * (size_t)E.nodes_per_tree + ts.root) * L::SIZE : nullptr;
    for (int a = 0; a < A; ++a) {
        n[(size_t)g * A + a] = nd ? ((const int*)(nd + L::OFF_N))[a] : 0;
        w[(size_t)g * A + a] = nd ? ((const double*)(nd + L::OFF_W))[a] : 0.0;
    }
    root_n[g] = ts.root_n; root_w[g] = ts.root_w;
    valid[g] = nd ? (unsigned short)(*(const unsigned*)(nd + L::OFF_META) & 0x1FFu) : 0;
}

__global__ void counters_kernel(EngineDev E, spx_counters* out) {
    __shared__ unsigned long long acc[9];
    if (threadIdx.x < 9) acc[threadIdx.x] = 0;
    __syncthreads();
    for (int g = threadIdx.x; g < E.cfg.n_games; g += blockDim.x) {
        const GameState& s = E.games[g];
        atomicAdd(&acc[0], s.cnt_sims); atomicAdd(&acc[1], s.cnt_evals); atomicAdd(&acc[2], s.cnt_term); atomicAdd(&acc[3], s.cnt_path);
        atomicAdd(&acc[4], s.cnt_moves); atomicAdd(&acc[5], s.cnt_games); atomicAdd(&acc[6], s.cnt_nodes); atomicAdd(&acc[7], s.cnt_err);
        atomicAdd(&acc[8], s.cnt_hits);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        out->sims = acc[0]; out->leaf_evals = acc[1]; out->terminal_sims = acc[2]; out->path_len_sum = acc[3];
        out->moves = acc[4]; out->games_finished = acc[5]; out->nodes_allocated = acc[6]; out->errors = acc[7];
        out->ticks = *E.ticks; out->records_dropped = *E.rec_dropped; out->cache_hits = acc[8];
    }
}

__global__ void pending_tree_kernel(EngineDev E, int* out) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < E.cfg.n_games) out[g] = E.games[g].pend_kind != PK_NONE ? E.games[g].pend_tree : -1;
}

__global__ void idle_kernel(EngineDev E, int* out) {
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    for (int g = threadIdx.x; g < E.cfg.n_games; g += blockDim.x)
        if (E.games[g].phase != PH_IDLE) any = 1;
    __syncthreads();
    if (threadIdx.x == 0) *out = any ? 0 : 1;
}

// ------------------------------------------------------------------------------------------------ env / hashnet kernels
template <int GAME>
__global__ void __launch_bounds__(256) env_step_kernel(long long n, ulonglong2* __restrict__ state, unsigned char* __restrict__ done,
                                const int* __restrict__ action, const signed char* __restrict__ player,
                                signed char* __restrict__ reward, unsigned short* __restrict__ valid,
                                signed char* __restrict__ status) {
    // grid-stride loop, 4 boards per thread per trip with all loads issued before any use (memory-level parallelism);
    // every load/store of a warp is one contiguous run: 512 B of state, 128 B of actions, 32 B of byte flags.
    constexpr int U = 4;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < n; base += stride * U) {
        ulonglong2 st[U];
        int a[U], d[U], pl[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = base + u * stride;
            if (i < n) { st[u] = state[i]; a[u] = action[i]; d[u] = done[i]; pl[u] = player[i]; }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = base + u * stride;
            if (i >= n) continue;
            int r = 0, code = SPX_ENV_OK;
            if (a[u] < 0) code = SPX_ENV_SKIPPED;
            else if (d[u]) code = SPX_ENV_GAME_OVER;
            else {
                u64 own = st[u].x, opp = st[u].y;
                int dn = 0;
                code = env_step<GAME>(own, opp, a[u], pl[u], r, dn);
                if (code == SPX_ENV_OK) { st[u].x = own; st[u].y = opp; state[i] = st[u]; done[i] = (unsigned char)dn; }
            }
            reward[i] = (signed char)r;
            valid[i] = (unsigned short)valid_mask<GAME>(st[u].x, st[u].y);
            status[i] = (signed char)code;
        }
    }
}

// Four CONSECUTIVE boards per thread, written for the DRAM roofline: the 64 bytes of state move as two 256-bit accesses (whole
// 32-byte sectors per lane), every byte-sized array with ONE vector access per thread (action int4, player/done/reward/status
// 4 x 8 bit, valid mask 4 x 16 bit), so a warp touches whole 128-byte lines; the rules are c4_step_fast / ttt_step_fast
// (Connect4: ~80 integer-pipe instructions per board; the generic restatement of env_step_kernel needs ~195, which made the
// first vector kernel ALU bound at 0.56 of the copy bandwidth).  Needs 32-byte aligned state and 16-byte aligned other arrays
// (the host wrapper checks and otherwise takes the scalar kernel).
template <int GAME>
__global__ void __launch_bounds__(256) env_step_quad_kernel(long long n4, ulonglong2* __restrict__ state, uchar4* __restrict__ done,
                                const int4* __restrict__ action, const char4* __restrict__ player, char4* __restrict__ reward,
                                ushort4* __restrict__ valid, char4* __restrict__ status) {
    __shared__ u64 s_lines[64];
    if (threadIdx.x < 64) s_lines[threadIdx.x] = GAME == SPX_GAME_CONNECT4 ? c4_lines_through(threadIdx.x) : ttt_lines_through(threadIdx.x);
    __syncthreads();
    // plain ld/st: .cs and L1::no_allocate variants measured the same (5.5 TB/s on 16 Mi boards)
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4; q += stride) {
        u64 own[4], opp[4];
        ulonglong2* sp = state + 4 * q;
        asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(own[0]), "=l"(opp[0]), "=l"(own[1]), "=l"(opp[1]) : "l"(sp) : "memory");
        asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(own[2]), "=l"(opp[2]), "=l"(own[3]), "=l"(opp[3]) : "l"(sp + 2) : "memory");
        int4 a4; unsigned d4w, p4w;
        asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(a4.x), "=r"(a4.y), "=r"(a4.z), "=r"(a4.w) : "l"(action + q) : "memory");
        asm volatile("ld.global.u32 %0, [%1];" : "=r"(d4w) : "l"(done + q) : "memory");
        asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(p4w) : "l"(player + q) : "memory");
        const int a[4] = {a4.x, a4.y, a4.z, a4.w};
        unsigned dn_w = 0, r_w = 0, code_w = 0;
        unsigned vm[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int r, dn, code;
            const int d = (d4w >> (8 * u)) & 0xFF;
            const int pl = (int)(signed char)((p4w >> (8 * u)) & 0xFF);
            step_fast<GAME>(own[u], opp[u], a[u], pl, d, s_lines, r, dn, code, vm[u]);
            dn_w |= (unsigned)dn << (8 * u);
            r_w |= (unsigned)r << (8 * u);
            code_w |= ((unsigned)code & 0xFFu) << (8 * u);
        }
        asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(sp), "l"(own[0]), "l"(opp[0]), "l"(own[1]), "l"(opp[1]) : "memory");
        asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" :: "l"(sp + 2), "l"(own[2]), "l"(opp[2]), "l"(own[3]), "l"(opp[3]) : "memory");
        asm volatile("st.global.u32 [%0], %1;" :: "l"(done + q), "r"(dn_w) : "memory");
        asm volatile("st.global.u32 [%0], %1;" :: "l"(reward + q), "r"(r_w) : "memory");
        asm volatile("st.global.v2.u32 [%0], {%1,%2};" :: "l"(valid + q), "r"(vm[0] | (vm[1] << 16)), "r"(vm[2] | (vm[3] << 16)) : "memory");
        asm volatile("st.global.u32 [%0], %1;" :: "l"(status + q), "r"(code_w) : "memory");
    }
}

template <int GAME>
__global__ void env_valid_kernel(long long n, const ulonglong2* __restrict__ state, unsigned short* __restrict__ valid) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        ulonglong2 st = state[i];
        valid[i] = (unsigned short)valid_mask<GAME>(st.x, st.y);
    }
}

__global__ void hashnet_kernel(int A, long long n, const u64* __restrict__ own, const u64* __restrict__ opp,
                               const unsigned char* __restrict__ needs, const unsigned char* __restrict__ net_id,
                               u64 seed0, u64 seed1, float* __restrict__ policy, float* __restrict__ value) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (needs && !needs[i]) return;
    const u64 seed = (net_id && net_id[i]) ? seed1 : seed0;
    const u64 k = sm64(own[i] ^ sm64(opp[i] + seed));
    unsigned r[SPX_MAX_ACTIONS], tot = 0;
    for (int a = 0; a < A; ++a) { r[a] = (unsigned)(sm64(k ^ (u64)(a + 1)) >> 48) + 1u; tot += r[a]; }
    const float ftot = (float)tot;
    for (int a = 0; a < A; ++a) policy[i * A + a] = __fdiv_rn((float)r[a], ftot);
    const float vv = __fsub_rn((float)(unsigned)(sm64(k ^ 0xFFULL) >> 48), 32768.0f);
    value[i] = __fdiv_rn(vv, 81920.0f);
}

// ------------------------------------------------------------------------------------------------ two-network leaf routing
// Head-to-head evaluation (compare_models / elo.py): tree 0 is evaluated by network 0, tree 1 by network 1.  The leaves of
// each network are packed to the front of their own dense batch (stable, by slot index) so that each tower only pays for
// its own rows; one CTA scans the whole batch (G <= 65536).
__global__ void __launch_bounds__(1024) partition_kernel(int n, const u64* __restrict__ own, const u64* __restrict__ opp,
                                                         const unsigned char* __restrict__ needs, const unsigned char* __restrict__ net_id,
                                                         u64* __restrict__ own2, u64* __restrict__ opp2, unsigned char* __restrict__ needs2,
                                                         int* __restrict__ map2) {
    __shared__ int warp_cnt[2][32];
    __shared__ int start[2];   // rows already written to each list before the current pass
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 2) start[tid] = 0;
    for (int i = tid; i < 2 * n; i += 1024) needs2[i] = 0;
    __syncthreads();
    for (int s = 0; s < n; s += 1024) {
        const int i = s + tid;
        const int which = (i < n && needs[i]) ? (net_id[i] ? 1 : 0) : -1;
        const unsigned m0 = __ballot_sync(0xffffffffu, which == 0), m1 = __ballot_sync(0xffffffffu, which == 1);
        const unsigned lt = (1u << lane) - 1u;
        if (lane == 0) { warp_cnt[0][warp] = __popc(m0); warp_cnt[1][warp] = __popc(m1); }
        __syncthreads();
        if (which >= 0) {
            int dst = start[which] + __popc((which ? m1 : m0) & lt);
            for (int w = 0; w < warp; ++w) dst += warp_cnt[which][w];
            own2[(size_t)which * n + dst] = own[i];
            opp2[(size_t)which * n + dst] = opp[i];
            needs2[(size_t)which * n + dst] = 1;
            map2[(size_t)which * n + dst] = i;
        }
        __syncthreads();
        if (tid < 2) { int tot = 0; for (int w = 0; w < 32; ++w) tot += warp_cnt[tid][w]; start[tid] += tot; }
        __syncthreads();
    }
}

__global__ void scatter_kernel(int n, int A, const unsigned char* __restrict__ needs2, const int* __restrict__ map2,
                               const float* __restrict__ pol2, const float* __restrict__ val2, float* __restrict__ policy,
                               float* __restrict__ value) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;   // row of the concatenated [2][n] dense batches
    if (j >= 2 * n || !needs2[j]) return;
    const int i = map2[j];
    for (int a = 0; a < A; ++a) policy[(size_t)i * A + a] = pol2[(size_t)j * A + a];
    value[i] = val2[j];
}

}  // namespace spx

// ================================================================================================== C ABI
using namespace spx;

// ---- device-to-device drain of the record ring, sorted by (game_index, tree, ply) (spx_drain_records_device)
__global__ void record_keys_kernel(const spx_record* __restrict__ ring, long long n, unsigned long long* __restrict__ keys, unsigned* __restrict__ vals) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        keys[i] = (ring[i].game_index << 8) | ((unsigned long long)(ring[i].tree & 1) << 7) | (unsigned long long)(ring[i].ply & 0x7F);
        vals[i] = (unsigned)i;
    }
}
__global__ void record_permute_kernel(const spx_record* __restrict__ ring, const unsigned* __restrict__ order, long long n, spx_record* __restrict__ out) {
    const uint4* s = reinterpret_cast<const uint4*>(ring);
    uint4* d = reinterpret_cast<uint4*>(out);
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < n * 5; w += (long long)gridDim.x * blockDim.x) {
        const long long rec = w / 5, part = w - rec * 5;
        d[w] = s[(long long)order[rec] * 5 + part];
    }
}

template <typename T>
static int drain(T* ring, unsigned long long* count, int64_t ring_cap, T* host_out, int64_t capacity, int64_t* n_out, cudaStream_t st) {
    unsigned long long n = 0;
    SPX_CUDA(cudaMemcpyAsync(&n, count, sizeof(n), cudaMemcpyDeviceToHost, st));
    SPX_CUDA(cudaStreamSynchronize(st));
    if ((int64_t)n > ring_cap) n = (unsigned long long)ring_cap;
    if ((int64_t)n > capacity) return set_err(SPX_E_OVERFLOW, "drain: host buffer smaller than the buffered entries%s", "");
    if (n) SPX_CUDA(cudaMemcpyAsync(host_out, ring, sizeof(T) * n, cudaMemcpyDeviceToHost, st));
    SPX_CUDA(cudaMemsetAsync(count, 0, sizeof(unsigned long long), st));
    SPX_CUDA(cudaStreamSynchronize(st));
    *n_out = (int64_t)n;
    return 0;
}

extern "C" {

const char* spx_last_error(void) { return g_err; }
int spx_version(void) { return 100; }
uint64_t spx_launch_count(void) { return g_launches.load(); }

static int grid_for(long long n, int block) {
    long long b = (n + block - 1) / block;
    const long long cap = 148LL * 8;  // persistent-style cap: 8 resident 256-thread CTAs on each of the 148 SMs
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// grid for a grid-stride kernel: exactly the CTAs that are resident at once (SM count x occupancy of this kernel), one wave
static int resident_grid(const void* kernel, long long n, int block) {
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (sms <= 0) sms = 148; }
    int per_sm = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, 0) != cudaSuccess || per_sm <= 0) per_sm = 4;
    const long long b = (n + block - 1) / block, cap = (long long)sms * per_sm;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

int spx_env_step(int32_t game, int64_t n, void* state, uint8_t* done, const int32_t* action, const int8_t* player,
                 int8_t* reward, uint16_t* valid, int8_t* status, void* stream) {
    if (n <= 0) return 0;
    if (!state || !done || !action || !player || !reward || !valid || !status) return set_err(SPX_E_ARG, "spx_env_step: null pointer%s", "");
    cudaStream_t st = (cudaStream_t)stream;
    if (game != SPX_GAME_CONNECT4 && game != SPX_GAME_TICTACTOE) return set_err(SPX_E_ARG, "spx_env_step: unknown game%s", "");
    const int block = 256;
    // vector path for the aligned bulk (4 boards per thread), scalar path for an unaligned call and for the last n % 4 boards
    const uintptr_t mis = (uintptr_t)state | (uintptr_t)done | (uintptr_t)action | (uintptr_t)player | (uintptr_t)reward | (uintptr_t)valid | (uintptr_t)status;
    const int64_t n4 = ((mis & 15) || ((uintptr_t)state & 31)) ? 0 : n / 4;
    if (n4 > 0) {
        if (game == SPX_GAME_CONNECT4)
            env_step_quad_kernel<SPX_GAME_CONNECT4><<<resident_grid((const void*)env_step_quad_kernel<SPX_GAME_CONNECT4>, n4, block), block, 0, st>>>(
                n4, (ulonglong2*)state, (uchar4*)done, (const int4*)action, (const char4*)player, (char4*)reward, (ushort4*)valid, (char4*)status);
        else
            env_step_quad_kernel<SPX_GAME_TICTACTOE><<<resident_grid((const void*)env_step_quad_kernel<SPX_GAME_TICTACTOE>, n4, block), block, 0, st>>>(
                n4, (ulonglong2*)state, (uchar4*)done, (const int4*)action, (const char4*)player, (char4*)reward, (ushort4*)valid, (char4*)status);
        count_launch();
    }
    const int64_t first = 4 * n4, rest = n - first;
    if (rest > 0) {
        const int grid = grid_for(rest, block);
        ulonglong2* s2 = (ulonglong2*)state + first;
        if (game == SPX_GAME_CONNECT4)
            env_step_kernel<SPX_GAME_CONNECT4><<<grid, block, 0, st>>>(rest, s2, done + first, action + first, (const signed char*)player + first, (signed char*)reward + first, valid + first, (signed char*)status + first);
        else
            env_step_kernel<SPX_GAME_TICTACTOE><<<grid, block, 0, st>>>(rest, s2, done + first, action + first, (const signed char*)player + first, (signed char*)reward + first, valid + first, (signed char*)status + first);
        count_launch();
    }
    SPX_CUDA(cudaGetLastError());
    return 0;
}

int spx_env_valid_moves(int32_t game, int64_t n, const void* state, uint16_t* valid, void* stream) {
    if (n <= 0) return 0;
    if (!state || !valid) return set_err(SPX_E_ARG, "spx_env_valid_moves: null pointer%s", "");
    cudaStream_t st = (cudaStream_t)stream;
    const int block = 256, grid = grid_for(n, block);
    if (game == SPX_GAME_CONNECT4) env_valid_kernel<SPX_GAME_CONNECT4><<<grid, block, 0, st>>>(n, (const ulonglong2*)state, valid);
    else if (game == SPX_GAME_TICTACTOE) env_valid_kernel<SPX_GAME_TICTACTOE><<<grid, block, 0, st>>>(n, (const ulonglong2*)state, valid);
    else return set_err(SPX_E_ARG, "spx_env_valid_moves: unknown game%s", "");
    count_launch();
    SPX_CUDA(cudaGetLastError());
    return 0;
}

int spx_hashnet_forward(int32_t game, int64_t n, const uint64_t* own, const uint64_t* opp, const uint8_t* needs_eval,
                        const uint8_t* net_id, uint64_t net_seed0, uint64_t net_seed1, float* policy, float* value, void* stream) {
    if (n <= 0) return 0;
    if (!own || !opp || !policy || !value) return set_err(SPX_E_ARG, "spx_hashnet_forward: null pointer%s", "");
    const int A = game == SPX_GAME_CONNECT4 ? 7 : 9;
    hashnet_kernel<<<(int)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(A, n, (const u64*)own, (const u64*)opp, needs_eval, net_id,
                                                                             net_seed0, net_seed1, policy, value);
    count_launch();
    SPX_CUDA(cudaGetLastError());
    return 0;
}

static int max_moves_of(int game) { return game == SPX_GAME_CONNECT4 ? 42 : 9; }
static int node_size_of(int game) { return game == SPX_GAME_CONNECT4 ? NodeLayout<SPX_GAME_CONNECT4>::SIZE : NodeLayout<SPX_GAME_TICTACTOE>::SIZE; }

int spx_create(const spx_config* cfg, spx_engine** out) {
    if (!cfg || !out) return set_err(SPX_E_ARG, "spx_create: null argument%s", "");
    if (cfg->game != SPX_GAME_CONNECT4 && cfg->game != SPX_GAME_TICTACTOE) return set_err(SPX_E_ARG, "spx_create: unknown game%s", "");
    if (cfg->n_games <= 0 || cfg->sims < 0 || cfg->slot_stride <= 0) return set_err(SPX_E_ARG, "spx_create: bad n_games/sims/slot_stride%s", "");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return set_err(SPX_E_CUDA, "spx_create: no CUDA device (there is no CPU fallback)%s", "");
    spx_engine* e = new (std::nothrow) spx_engine();
    if (!e) return set_err(SPX_E_ARG, "spx_create: out of host memory%s", "");
    memset(e, 0, sizeof(*e));
    EngineDev& d = e->d;
    d.cfg = *cfg;
    if (d.cfg.max_sims_per_tick < 1) d.cfg.max_sims_per_tick = 1;
    if (cfg->search_threads < 0 || cfg->search_threads > 16) return set_err(SPX_E_ARG, "spx_create: search_threads must be in [0, 16]%s", "");
    d.K = cfg->search_threads > 1 ? cfg->search_threads : 1;
    if (d.cfg.record_capacity < 1) d.cfg.record_capacity = 1;
    if (d.cfg.result_capacity < 1) d.cfg.result_capacity = 1;
    const int mm = max_moves_of(cfg->game);
    d.nodes_per_tree = cfg->nodes_per_tree > 0 ? cfg->nodes_per_tree : (cfg->sims + 1) * ((mm + 1) / 2) + mm + 2;
    if ((long long)d.nodes_per_tree >= (1LL << 28)) return set_err(SPX_E_ARG, "spx_create: nodes_per_tree too large%s", "");
    SPX_CUDA(cudaGetDevice(&e->device));
    const size_t G = (size_t)cfg->n_games;
    int64_t bytes = 0;
#define SPX_ALLOC(ptr, type, count)                                   \
    do {                                                              \
        size_t _b = sizeof(type) * (size_t)(count);                   \
        SPX_CUDA(cudaMalloc((void**)&(ptr), _b ? _b : 16));           \
        SPX_CUDA(cudaMemset((void*)(ptr), 0, _b ? _b : 16));          \
        bytes += (int64_t)_b;                                         \
    } while (0)
    SPX_ALLOC(d.games, GameState, G);
    {
        size_t pb = G * 2 * (size_t)d.nodes_per_tree * node_size_of(cfg->game);
        SPX_CUDA(cudaMalloc((void**)&d.pool, pb));
        bytes += (int64_t)pb;
    }
    SPX_ALLOC(d.paths, unsigned, G * SPX_MAX_PATH);
    SPX_ALLOC(d.noise, double, G * SPX_MAX_ACTIONS);
    SPX_ALLOC(d.temp_rec, spx_record, G * 2 * SPX_MAX_OWN_MOVES);
    if (cfg->move_log) SPX_ALLOC(d.mlog, spx_move_log, G * SPX_MAX_PLIES);
    SPX_ALLOC(d.rec_ring, spx_record, d.cfg.record_capacity);
    SPX_ALLOC(d.res_ring, spx_result, d.cfg.result_capacity);
    SPX_ALLOC(d.rec_count, unsigned long long, 1);
    SPX_ALLOC(d.res_count, unsigned long long, 1);
    SPX_ALLOC(d.rec_dropped, unsigned long long, 1);
    SPX_ALLOC(d.ticks, unsigned long long, 2);   // [0] ticks run, [1] pass tickets of the running work-conserving launch (spx_tick_fused_balanced)
    SPX_ALLOC(d.leaf_own, u64, G * d.K);
    SPX_ALLOC(d.leaf_opp, u64, G * d.K);
    SPX_ALLOC(d.needs_eval, unsigned char, G * d.K);
    SPX_ALLOC(d.net_id, unsigned char, G * d.K);
    if (d.K > 1) {   // threaded search: worker contexts, their select paths, MCNode.virtual_loss of every node's children
        SPX_ALLOC(d.workers, Worker, G * d.K);
        SPX_ALLOC(d.wpaths, unsigned, G * d.K * SPX_MAX_PATH);
        SPX_ALLOC(d.vlpool, unsigned short, G * 2 * (size_t)d.nodes_per_tree * (cfg->game == SPX_GAME_TICTACTOE ? 16 : 8));
    }
    SPX_ALLOC(d.ext_action, int, G);
    SPX_ALLOC(d.own_action, int, 2 * G);
    SPX_ALLOC(d.sqrt_table, double, sf::SQRT_TABLE);
    sqrt_table_kernel<<<sf::SQRT_TABLE / 256, 256>>>(const_cast<double*>(d.sqrt_table));
    count_launch();
    if (cfg->eval_cache_log2) {   // per slot a direct-mapped table of 64-byte entries (spx_advance.cuh: evaluation cache)
        if (cfg->eval_cache_log2 < 6 || cfg->eval_cache_log2 > 20) return set_err(SPX_E_ARG, "spx_create: eval_cache_log2 must be 0 (off) or in [6, 20]%s", "");
        if (d.K > 1) return set_err(SPX_E_ARG, "spx_create: the evaluation cache serves the sequential search (search_threads <= 1)%s", "");
        SPX_ALLOC(e->ecache, uint4, (G << cfg->eval_cache_log2) * 4);
    }
#undef SPX_ALLOC
    e->bytes = bytes;
    *out = e;
    return spx_reset(e, nullptr);
}

int spx_destroy(spx_engine* e) {
    if (!e) return 0;
    EngineDev& d = e->d;
    void* ptrs[] = {d.games, d.pool, d.paths, d.noise, d.temp_rec, d.mlog, d.rec_ring, d.res_ring, d.rec_count, d.res_count,
                    d.rec_dropped, d.ticks, d.leaf_own, d.leaf_opp, d.needs_eval, d.net_id, d.ext_action, d.own_action, d.workers, d.wpaths, d.vlpool, e->ecache, (void*)d.sqrt_table};
    for (void* p : ptrs) if (p) cudaFree(p);
    delete e;
    return 0;
}

int spx_reset(spx_engine* e, void* stream) {
    if (!e) return set_err(SPX_E_ARG, "spx_reset: null engine%s", "");
    reset_kernel<<<(e->d.cfg.n_games + 127) / 128, 128, 0, (cudaStream_t)stream>>>(e->d);
    count_launch();
    SPX_CUDA(cudaGetLastError());
    return 0;
}

int spx_set_noise_table(spx_engine* e, const double* table, int64_t first_game_index, int64_t n_table_games, int32_t table_moves) {
    if (!e) return set_err(SPX_E_ARG, "spx_set_noise_table: null engine%s", "");
    e->d.noise_table = table; e->d.table_first = first_game_index; e->d.table_games = n_table_games; e->d.table_moves = table_moves;
    return 0;
}

int spx_advance(spx_engine* e, const float* policy, const float* value, void* stream) {
    if (!e) return set_err(SPX_E_ARG, "spx_advance: null engine%s", "");
    const int block = 128, grid = (e->d.cfg.n_games * 32 + block - 1) / block;
    if (e->d.K > 1) {   // threaded search: K leaf slots per game
        if (e->d.cfg.game == SPX_GAME_CONNECT4) advance_kernel_threaded<SPX_GAME_CONNECT4><<<grid, block, 0, (cudaStream_t)stream>>>(e->d, policy, value);
        else advance_kernel_threaded<SPX_GAME_TICTACTOE><<<grid, block, 0, (cudaStream_t)stream>>>(e->d, policy, value);
    } else if (e->ecache && e->cache_ver[0] && (!e->d.cfg.two_nets || e->cache_ver[1])) {
        // evaluation cache: the caller has said which weights versions its policy / value inputs come from
        EngineDev d = e->d;
        d.ecache = e->ecache;
        d.ecache_log2 = (unsigned)d.cfg.eval_cache_log2;
        d.ecache_tag[0] = (e->cache_ver[0] << 8) | 1u;
        d.ecache_tag[1] = (e->cache_ver[1] << 8) | 3u;
        if (d.cfg.game == SPX_GAME_CONNECT4) advance_kernel_cached<SPX_GAME_CONNECT4><<<grid, block, 0, (cudaStream_t)stream>>>(d, policy, value);
        else advance_kernel_cached<SPX_GAME_TICTACTOE><<<grid, block, 0, (cudaStream_t)stream>>>(d, policy, value);
    } else if (e->d.cfg.game == SPX_GAME_CONNECT4) advance_kernel<SPX_GAME_CONNECT4><<<grid, block, 0, (cudaStream_t)stream>>>(e->d, policy, value);
    else advance_kernel<SPX_GAME_TICTACTOE><<<grid, block, 0, (cudaStream_t)stream>>>(e->d, policy, value);
    count_launch();
    SPX_CUDA(cudaGetLastError());
    return 0;
}

int spx_softf64_selftest(uint64_t n, uint64_t seed, uint64_t* mismatches_out, void* stream) {
    if (!mismatches_out) return set_err(SPX_E_ARG, "spx_softf64_selftest: null output%s", "");
    cudaStream_t st = (cudaStream_t)stream;
    double* table = nullptr;
    unsigned long long* mism = nullptr;
    SPX_CUDA(cudaMalloc((void**)&table, sizeof(double) * sf::SQRT_TABLE));
    SPX_CUDA(cudaMalloc((void**)&mism, 8 * sizeof(unsigned long long)));
    SPX_CUDA(cudaMemsetAsync(mism, 0, 8 * sizeof(unsigned long long), st));
    sqrt_table_kernel<<<sf::SQRT_TABLE / 256, 256, 0, st>>>(table);
    softf64_selftest_kernel<<<1184, 256, 0, st>>>((unsigned long long)n, (unsigned long long)seed, mism, table);
    count_launch(); count_launch();
    SPX_CUDA(cudaGetLastError());
    SPX_CUDA(cudaMemcpyAsync(mismatches_out, mism, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    SPX_CUDA(cudaStreamSynchronize(st));
    cudaFree(table); cudaFree(mism);
    return 0;
}

int spx_set_eval_cache_versions(spx_engine* e, uint32_t version0, uint32_t version1) {
    if (!e) return set_err(SPX_E_ARG, "spx_set_eval_cache_versions: null engine%s", "");
    e->cache_ver[0] = version0 & 0xFFFFFFu;
    e->cache_ver[1] = version1 & 0xFFFFFFu;
    return 0;
}

int spx_event_create(void** ev_out) {
    if (!ev_out) return set_err(SPX_E_ARG, "spx_event_create: null%s", "");
    cudaEvent_t e;
    SPX_CUDA(cudaEventCreate(&e));
    *ev_out = (void*)e;
    return 0;
}
int spx_event_destroy(void* ev) { if (ev) cudaEventDestroy((cudaEvent_t)ev); return 0; }
int spx_event_elapsed_ms(void* ev_start, void* ev_end, float* ms_out) {
    if (!ev_start || !ev_end || !ms_out) return set_err(SPX_E_ARG, "spx_event_elapsed_ms: null%s", "");
    SPX_CUDA(cudaEventSynchronize((cudaEvent_t)ev_end));
    SPX_CUDA(cudaEventElapsedTime(ms_out, (cudaEvent_t)ev_start, (cudaEvent_t)ev_end));
    return 0;
}

int spx_advance_timed(spx_engine* e, const float* policy, const float* value, void* stream, void* ev_start, void* ev_end) {
    if (ev_start) SPX_CUDA(cudaEventRecord((cudaEvent_t)ev_start, (cudaStream_t)stream));
    int rc = spx_advance(e, policy, value, stream);
    if (rc) return rc;
    if (ev_end) SPX_CUDA(cudaEventRecord((cudaEvent_t)ev_end, (cudaStream_t)stream));
    return 0;
}

int spx_leaf_batch(spx_engine* e, uint64_t** own, uint64_t** opp, uint8_t** needs_eval, uint8_t** net_id) {
    if (!e) return set_err(SPX_E_ARG, "spx_leaf_batch: null engine%s", "");
    if (own) *own = (uint64_t*)e->d.leaf_own;
    if (opp) *opp = (uint64_t*)e->d.leaf_opp;
    if (needs_eval) *needs_eval = e->d.needs_eval;
    if (net_id) *net_id = e->d.net_id;
    return 0;
}

int spx_root_stats(spx_engine* e, int32_t tree, int32_t* n, double* w, int32_t* root_n, double* root_w, uint16_t* valid, void* stream) {
    if (!e || tree < 0 || tree > 1 || !n || !w || !root_n || !root_w || !valid) return set_err(SPX_E_ARG, "spx_root_stats: bad argument%s", "");
    const int grid = (e->d.cfg.n_games + 127) / 128;
    if (e->d.cfg.game == SPX_GAME_CONNECT4) root_stats_kernel<SPX_GAME_CONNECT4><<<grid, 128, 0, (cudaStream_t)stream>>>(e->d, tree, n, w, root_n, root_w, valid);
    else root_stats_kernel<SPX_GAME_TICTACTOE><<<grid, 128, 0, (cudaStream_t)stream>>>(e->d, tree, n, w, root_n, root_w, valid);
    count_launch();
    SPX_CUDA(cudaGetLastError());
    return 0;
}

int spx_drain_records(spx_engine* e, spx_record* host_out, int64_t capacity, int64_t* n_out, void* stream) {
    if (!e || !host_out || !n_out) return set_err(SPX_E_ARG, "spx_drain_records: bad argument%s", "");
    return drain<spx_record>(e->d.rec_ring, e->d.rec_count, e->d.cfg.record_capacity, host_out, capacity, n_out, (cudaStream_t)stream);
}
int spx_drain_records_device(spx_engine* e, spx_record* dev_out, int64_t capacity, int64_t* n_out, void* stream) {
    if (!e || !dev_out || !n_out) return set_err(SPX_E_ARG, "spx_drain_records_device: bad argument%s", "");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned long long n = 0;
    SPX_CUDA(cudaMemcpyAsync(&n, e->d.rec_count, sizeof(n), cudaMemcpyDeviceToHost, st));
    SPX_CUDA(cudaStreamSynchronize(st));
    if ((int64_t)n > e->d.cfg.record_capacity) n = (unsigned long long)e->d.cfg.record_capacity;
    if ((int64_t)n > capacity) return set_err(SPX_E_OVERFLOW, "spx_drain_records_device: output buffer smaller than the buffered records%s", "");
    *n_out = (int64_t)n;
    if (n == 0) return 0;
    // scratch: keys in/out, order in/out, CUB temp storage -- one allocation per drain (a per-epoch operation)
    size_t temp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const unsigned*)nullptr, (unsigned*)nullptr, (int)n);
    const size_t kb = (sizeof(unsigned long long) * n + 255) & ~(size_t)255, vb = (sizeof(unsigned) * n + 255) & ~(size_t)255;
    unsigned char* scratch = nullptr;
    SPX_CUDA(cudaMalloc((void**)&scratch, 2 * kb + 2 * vb + temp_bytes));
    unsigned long long *k0 = (unsigned long long*)scratch, *k1 = (unsigned long long*)(scratch + kb);
    unsigned *v0 = (unsigned*)(scratch + 2 * kb), *v1 = (unsigned*)(scratch + 2 * kb + vb);
    const int grid = grid_for((long long)n, 256);
    record_keys_kernel<<<grid, 256, 0, st>>>(e->d.rec_ring, (long long)n, k0, v0);
    count_launch();
    cudaError_t ce = cub::DeviceRadixSort::SortPairs(scratch + 2 * kb + 2 * vb, temp_bytes, k0, k1, v0, v1, (int)n, 0, 64, st);
    if (ce == cudaSuccess) {
        record_permute_kernel<<<grid_for((long long)n * 5, 256), 256, 0, st>>>(e->d.rec_ring, v1, (long long)n, dev_out);
        count_launch();
        ce = cudaGetLastError();
    }
    if (ce == cudaSuccess) ce = cudaMemsetAsync(e->d.rec_count, 0, sizeof(unsigned long long), st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    cudaFree(scratch);
    if (ce != cudaSuccess) return set_err(SPX_E_CUDA, "spx_drain_records_device: %s", cudaGetErrorString(ce));
    return 0;
}
int spx_drain_results(spx_engine* e, spx_result* host_out, int64_t capacity, int64_t* n_out, void* stream) {
    if (!e || !host_out || !n_out) return set_err(SPX_E_ARG, "spx_drain_results: bad argument%s", "");
    // results beyond the ring's capacity were not written (advance_game keeps counting): that is lost data, not a statistic
    unsigned long long pending = 0;
    SPX_CUDA(cudaMemcpyAsync(&pending, e->d.res_count, sizeof(pending), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    SPX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    const int rc = drain<spx_result>(e->d.res_ring, e->d.res_count, e->d.cfg.result_capacity, host_out, capacity, n_out, (cudaStream_t)stream);
    if (rc == 0 && (int64_t)pending > e->d.cfg.result_capacity)
        return set_err(SPX_E_OVERFLOW, "spx_drain_results: game results were lost, the result ring overflowed between two drains "
                                       "(drain more often or raise result_capacity)%s", "");
    return rc;
}

int spx_read_move_log(spx_engine* e, int32_t slot, spx_move_log* host_out, int32_t capacity, int32_t* n_out, void* stream) {
    if (!e || !host_out || !n_out || slot < 0 || slot >= e->d.cfg.n_games) return set_err(SPX_E_ARG, "spx_read_move_log: bad argument%s", "");
    if (!e->d.mlog) return set_err(SPX_E_STATE, "spx_read_move_log: engine created with move_log = 0%s", "");
    GameState s;
    SPX_CUDA(cudaMemcpyAsync(&s, e->d.games + slot, sizeof(s), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    SPX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    int n = s.n_moves_logged < capacity ? s.n_moves_logged : capacity;
    if (n) SPX_CUDA(cudaMemcpyAsync(host_out, e->d.mlog + (size_t)slot * SPX_MAX_PLIES, sizeof(spx_move_log) * n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    SPX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *n_out = n;
    return 0;
}

int spx_counters_read(spx_engine* e, spx_counters* host_out, void* stream) {
    if (!e || !host_out) return set_err(SPX_E_ARG, "spx_counters_read: bad argument%s", "");
    spx_counters* dtmp = nullptr;
    SPX_CUDA(cudaMalloc((void**)&dtmp, sizeof(spx_counters)));
    counters_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(e->d, dtmp);
    count_launch();
    cudaError_t err = cudaMemcpyAsync(host_out, dtmp, sizeof(spx_counters), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(dtmp);
    if (err != cudaSuccess) return set_err(SPX_E_CUDA, "spx_counters_read: %s", cudaGetErrorString(err));
    return 0;
}

int spx_all_idle(spx_engine* e, int32_t* idle_out, void* stream) {
    if (!e || !idle_out) return set_err(SPX_E_ARG, "spx_all_idle: bad argument%s", "");
    int* dtmp = nullptr;
    SPX_CUDA(cudaMalloc((void**)&dtmp, sizeof(int)));
    idle_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(e->d, dtmp);
    count_launch();
    cudaError_t err = cudaMemcpyAsync(idle_out, dtmp, sizeof(int), cudaMemcpyDeviceToHost, (cudaStream_t)stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(dtmp);
    if (err != cudaSuccess) return set_err(SPX_E_CUDA, "spx_all_idle: %s", cudaGetErrorString(err));
    return 0;
}

int64_t spx_device_bytes(spx_engine* e) { return e ? e->bytes : 0; }

int spx_partition_leaves(int64_t n, const uint64_t* own, const uint64_t* opp, const uint8_t* needs_eval, const uint8_t* net_id,
                         uint64_t* own2, uint64_t* opp2, uint8_t* needs2, int32_t* map2, void* stream) {
    if (n <= 0) return 0;
    if (n > 65536 || !own || !opp || !needs_eval || !net_id || !own2 || !opp2 || !needs2 || !map2) return set_err(SPX_E_ARG, "spx_partition_leaves: bad argument%s", "");
    partition_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>((int)n, (const u64*)own, (const u64*)opp, needs_eval, net_id, (u64*)own2, (u64*)opp2, needs2, map2);
    count_launch();
    SPX_CUDA(cudaGetLastError());
    return 0;
}

int spx_scatter_outputs(int64_t n, int32_t n_actions, const uint8_t* needs2, const int32_t* map2, const float* policy2, const float* value2,
                        float* policy, float* value, void* stream) {
    if (n <= 0) return 0;
    if (!needs2 || !map2 || !policy2 || !value2 || !policy || !value) return set_err(SPX_E_ARG, "spx_scatter_outputs: null pointer%s", "");
    scatter_kernel<<<(int)((2 * n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((int)n, n_actions, needs2, map2, policy2, value2, policy, value);
    count_launch();
    SPX_CUDA(cudaGetLastError());
    return 0;
}

int spx_restart(spx_engine* e, int64_t slot_offset, int64_t games_target, void* stream) {
    if (!e) return set_err(SPX_E_ARG, "spx_restart: null engine%s", "");
    e->d.cfg.slot_offset = slot_offset;
    e->d.cfg.games_target = games_target;
    return spx_reset(e, stream);
}

int spx_set_sims(spx_engine* e, int32_t sims) {
    if (!e) return set_err(SPX_E_ARG, "spx_set_sims: null engine%s", "");
    const int mm = e->d.cfg.game == SPX_GAME_TICTACTOE ? 9 : 42;
    // the node pool was sized for the simulations per move the engine was created with (it can never overflow)
    if (sims < 1 || (long long)(sims + 1) * ((mm + 1) / 2) + mm + 2 > (long long)e->d.nodes_per_tree)
        return set_err(SPX_E_ARG, "spx_set_sims: sims must be in [1, the value the node pool was sized for]%s", "");
    e->d.cfg.sims = sims;   // read by the next launch (the config travels by value); searches in progress simply run on to the new count
    return 0;
}

__global__ void set_external_kernel(EngineDev E, const int* actions) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g < E.cfg.n_games && actions[g] >= 0) E.ext_action[g] = actions[g];
}
__global__ void slot_status_kernel(EngineDev E, int* out) {
    int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= E.cfg.n_games) return;
    const GameState& s = E.games[g];
    const bool waiting = s.phase == PH_SEARCH && s.mover_tree == 1 && E.cfg.opponent_kind == SPX_OPP_EXTERNAL && E.ext_action[g] < 0;
    out[6 * g + 0] = s.phase == PH_IDLE ? 2 : (waiting ? 1 : 0);
    out[6 * g + 1] = s.ply;
    out[6 * g + 2] = E.own_action[2 * g];
    out[6 * g + 3] = E.own_action[2 * g + 1];
    out[6 * g + 4] = (int)s.cnt_games;
    out[6 * g + 5] = s.swap;
}

int spx_set_external_actions(spx_engine* e, const int32_t* actions, void* stream) {
    if (!e || !actions) return set_err(SPX_E_ARG, "spx_set_external_actions: bad argument%s", "");
    if (e->d.cfg.opponent_kind != SPX_OPP_EXTERNAL) return set_err(SPX_E_STATE, "spx_set_external_actions: engine not created with SPX_OPP_EXTERNAL%s", "");
    set_external_kernel<<<(e->d.cfg.n_games + 127) / 128, 128, 0, (cudaStream_t)stream>>>(e->d, actions);
    count_launch();
    SPX_CUDA(cudaGetLastError());
    return 0;
}

int spx_slot_status(spx_engine* e, int32_t* status_out, void* stream) {
    if (!e || !status_out) return set_err(SPX_E_ARG, "spx_slot_status: bad argument%s", "");
    slot_status_kernel<<<(e->d.cfg.n_games + 127) / 128, 128, 0, (cudaStream_t)stream>>>(e->d, status_out);
    count_launch();
    SPX_CUDA(cudaGetLastError());
    return 0;
}

int spx_pending_tree(spx_engine* e, int32_t* tree_out, void* stream) {
    if (!e || !tree_out) return set_err(SPX_E_ARG, "spx_pending_tree: bad argument%s", "");
    pending_tree_kernel<<<(e->d.cfg.n_games + 127) / 128, 128, 0, (cudaStream_t)stream>>>(e->d, tree_out);
    count_launch();
    SPX_CUDA(cudaGetLastError());
    return 0;
}

}  // extern "C"

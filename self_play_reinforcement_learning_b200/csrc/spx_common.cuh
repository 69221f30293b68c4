// spx_common.cuh -- device-side building blocks shared by the env and search kernels.
//   * counter-based RNG stream (twin of oracle/spec.py; integer-only, so CPU == GPU bit for bit)
//   * bitboard game rules for Connect4 (7x6) and TicTacToe (3x3), following
//     games/connect4/connect4env.py:29-48,72-92 and games/tictactoe/tictactoe_env.py:23-45,62-82
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spx.h"

typedef unsigned long long u64;

namespace spx {

enum { PURPOSE_TIE = 0, PURPOSE_GAMMA = 1, PURPOSE_ACTION = 2, PURPOSE_OPPONENT = 3 };

__host__ __device__ __forceinline__ u64 sm64(u64 x) {
    x += 0x9E3779B97F4A7C15ULL;
    u64 z = x;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

// prefix of the key that does not depend on (sim, depth, idx): hoisted out of the select loop
__host__ __device__ __forceinline__ u64 rng_prefix(u64 seed, u64 game_uid, int tree, int purpose, int ply) {
    u64 h = sm64(seed);
    h = sm64(h ^ game_uid);
    h = sm64(h ^ ((u64)(tree & 0xFF) | ((u64)(purpose & 0xFF) << 8) | ((u64)(ply & 0xFFFF) << 16)));
    return h;
}
__host__ __device__ __forceinline__ double rng_uniform_from(u64 prefix, uint32_t sim, uint32_t depth, u64 idx) {
    u64 h = sm64(prefix ^ ((u64)sim | ((u64)depth << 32)));
    h = sm64(h ^ idx);
    return (double)(h >> 11) * (1.0 / 9007199254740992.0);
}
// the same draw as its 53-bit integer k: the uniform is k * 2^-53
__host__ __device__ __forceinline__ u64 rng_uniform_bits(u64 prefix, uint32_t sim, uint32_t depth, u64 idx) {
    u64 h = sm64(prefix ^ ((u64)sim | ((u64)depth << 32)));
    h = sm64(h ^ idx);
    return h >> 11;
}

// ------------------------------------------------------------------------------------------------
template <int GAME> struct Rules;

template <> struct Rules<SPX_GAME_CONNECT4> {
    static constexpr int W = 7, H = 6, A = 7, STRIDE = 7, WIN = 4, CELLS = 42;
    static constexpr u64 BOARD = 0x0000FDFBF7EFDFBFULL;  // 6 playable bits in each of 7 columns
    __host__ __device__ static __forceinline__ int height(u64 occ, int c) { return __builtin_popcountll((occ >> (7 * c)) & 0x3FULL); }
};
template <> struct Rules<SPX_GAME_TICTACTOE> {
    static constexpr int W = 3, H = 3, A = 9, STRIDE = 3, WIN = 3, CELLS = 9;
    static constexpr u64 BOARD = 0x1FFULL;
};

#ifdef __CUDA_ARCH__
#define SPX_POPC(x) __popcll(x)
#else
#define SPX_POPC(x) __builtin_popcountll(x)
#endif

// Does `b` hold a run of >= WIN along the full line through cell (x,y) in direction (dx,dy)?
// This is functools.reduce(_calc_win_in_a_row, line*player, 0) == WIN (connect4env.py:80-92): the
// fold saturates once any run of WIN appears anywhere on that line, not only through (x,y).
template <int GAME>
__host__ __device__ __forceinline__ bool line_has_run(u64 b, int x, int y, int dx, int dy) {
    typedef Rules<GAME> R;
    // walk back to the first cell of the line
    int back = 0;
    {
        int bx = dx > 0 ? x : (dx < 0 ? R::W - 1 - x : 99);
        int by = dy > 0 ? y : (dy < 0 ? R::H - 1 - y : 99);
        back = bx < by ? bx : by;
    }
    int cx = x - back * dx, cy = y - back * dy;
    unsigned bits = 0;
    int len = 0;
#pragma unroll
    for (int i = 0; i < 7; ++i) {
        if (cx >= 0 && cx < R::W && cy >= 0 && cy < R::H) {
            bits |= (unsigned)((b >> (cx * R::STRIDE + cy)) & 1ULL) << len;
            ++len;
            cx += dx;
            cy += dy;
        }
    }
    unsigned m = bits & (bits >> 1) & (bits >> 2);
    if (R::WIN == 4) m &= (bits >> 3);
    return m != 0;
}

// get_reward(action, player): 1 iff one of the four lines through the last cell holds a run of WIN
// of the mover's pieces (connect4env.py:72-83, tictactoe_env.py:62-74)
template <int GAME>
__host__ __device__ __forceinline__ int reward_at(u64 mover_bits, int x, int y) {
    return (line_has_run<GAME>(mover_bits, x, y, 1, 0) ||   // board[:, y]
            line_has_run<GAME>(mover_bits, x, y, 0, 1) ||   // board[x, :]
            line_has_run<GAME>(mover_bits, x, y, 1, 1) ||   // np.diagonal(board, y - x)
            line_has_run<GAME>(mover_bits, x, y, -1, 1))    // np.diagonal(np.flipud(board), ...)
               ? 1 : 0;
}

// Connect4 fast path: the same four full-line folds, as shift-AND run detection on masked bitboards.
// (bit = col*7+row; the sentinel bit 6 of every column keeps shifted runs from leaking across columns.)
template <>
__host__ __device__ __forceinline__ int reward_at<SPX_GAME_CONNECT4>(u64 m, int x, int y) {
    const u64 v = m & (0x3FULL << (7 * x));                                  // board[x, :]
    u64 hit = v & (v >> 1) & (v >> 2) & (v >> 3);
    const u64 h = m & (0x0000040810204081ULL << y);                          // board[:, y]
    hit |= h & (h >> 7) & (h >> 14) & (h >> 21);
    u64 d1 = 0, d2 = 0;                                                      // the two diagonals through (x, y)
#pragma unroll
    for (int c = 0; c < 7; ++c) {
        const int r1 = c - (x - y), r2 = (x + y) - c;
        if (r1 >= 0 && r1 < 6) d1 |= 1ULL << (7 * c + r1);
        if (r2 >= 0 && r2 < 6) d2 |= 1ULL << (7 * c + r2);
    }
    d1 &= m; d2 &= m;
    hit |= d1 & (d1 >> 8) & (d1 >> 16) & (d1 >> 24);
    hit |= d2 & (d2 >> 6) & (d2 >> 12) & (d2 >> 18);
    return hit != 0 ? 1 : 0;
}

// Union of the four full lines (column, row, both diagonals) through Connect4 cell `bitpos` = 7*col + row; 0 for sentinel bits.
// Run detection on `mover & c4_lines_through(cell)` equals the four per-line folds of get_reward: a line parallel to direction d
// that does not pass through the cell meets the other three lines in at most three cells, so it cannot hold a run of four.
__host__ __device__ inline u64 c4_lines_through(int bitpos) {
    const int x = bitpos / 7, y = bitpos % 7;
    if (x > 6 || y > 5) return 0;
    u64 m = 0;
    for (int c = 0; c < 7; ++c)
        for (int r = 0; r < 6; ++r)
            if (c == x || r == y || c - r == x - y || c + r == x + y) m |= 1ULL << (7 * c + r);
    return m;
}

// Connect4 env.step for the batched kernel, branch-free and sized for the integer pipe (the scalar restatement above costs
// ~195 ALU instructions per board and made the kernel ALU bound):  * the four line folds run on ONE masked board
// (`lines` = shared-memory copy of c4_lines_through, 49 entries)  * the column fold runs on the 6-bit column in 32-bit registers
// * heights < 6 of all seven columns is one 64-bit multiply that gathers the top-row bits.   Same results as env_step<CONNECT4>
// + valid_mask<CONNECT4> bit for bit (tests/test_env_gpu.py holds the two against each other); action >= 7 reports VALUE_ERROR.
__device__ __forceinline__ void c4_step_fast(u64& own, u64& opp, int action, int player, int done_in, const u64* __restrict__ lines,
                                             int& reward, int& done_out, int& code, unsigned& valid) {
    const u64 occ = own | opp;
    const bool in_range = (unsigned)action < 7u;
    const int sh = in_range ? 7 * action : 0;
    const unsigned col = (unsigned)(occ >> sh) & 0x3Fu;
    const int y = __popc(col);                                      // heights[action]
    const bool step = in_range && !done_in && y < 6;
    code = action < 0 ? SPX_ENV_SKIPPED : (done_in ? SPX_ENV_GAME_OVER : (step ? SPX_ENV_OK : SPX_ENV_VALUE_ERROR));
    const int cell = sh + y;
    // the new piece as two 32-bit halves: shl.b32 clamps (amount >= 32 gives 0), so "no step" is simply amount 64
    const unsigned amt = step ? (unsigned)cell : 64u;
    unsigned bit_lo, bit_hi;
    asm("shl.b32 %0, 1, %1;" : "=r"(bit_lo) : "r"(amt));
    asm("shl.b32 %0, 1, %1;" : "=r"(bit_hi) : "r"(amt - 32u));
    const u64 bit = ((u64)bit_hi << 32) | bit_lo;
    const u64 mine = player > 0 ? ~0ULL : 0ULL;                     // select by mask: folds into three-input logic ops
    own |= bit & mine;
    opp |= bit & ~mine;
    const u64 m = (own & mine) | (opp & ~mine);
    const unsigned cm = (unsigned)(m >> sh) & 0x3Fu;                // board[x, :]
    unsigned v = cm & (cm >> 1);
    v &= v >> 2;
    const u64 mm = m & lines[cell];
    u64 h = mm & (mm >> 7);   h &= h >> 14;                         // board[:, y]
    u64 d1 = mm & (mm >> 8);  d1 &= d1 >> 16;                       // np.diagonal(board, y - x)
    u64 d2 = mm & (mm >> 6);  d2 &= d2 >> 12;                       // np.diagonal(np.flipud(board), ...)
    const u64 hit = h | d1 | d2;
    reward = (step && (v | (unsigned)hit | (unsigned)(hit >> 32))) ? 1 : 0;
    const u64 occ2 = occ | bit;
    const u64 open_top = ~occ2 & 0x0000810204081020ULL;             // bit 7c+5 clear <=> heights[c] < 6
    valid = (unsigned)((open_top * 0x0000001041041041ULL) >> 41) & 0x7Fu;   // bit 7c+5 -> bit c (no two partial products collide)
    done_out = step ? (int)((reward != 0) | (__popcll(occ2) == 42)) : done_in;
}

// TicTacToe: the (up to) four lines of three through cell `a` = 3x+y as four 16-bit fields (row board[:, y], column board[x, :],
// main diagonal if x == y, anti-diagonal if x + y == 2); an absent or shorter line is 0xFFFF, which no 9-bit board contains.
__host__ __device__ inline u64 ttt_lines_through(int a) {
    if (a < 0 || a > 8) return ~0ULL;
    const int x = a / 3, y = a % 3;
    const u64 row = 0x49ULL << y, colm = 0x7ULL << (3 * x);
    const u64 dg = (x == y) ? 0x111ULL : 0xFFFFULL, ad = (x + y == 2) ? 0x54ULL : 0xFFFFULL;
    return row | (colm << 16) | (dg << 32) | (ad << 48);
}

// TicTacToe env.step for the batched kernel (same results as env_step<TICTACTOE> + valid_mask<TICTACTOE>, bit for bit):
// occupied target = silent no-op with the reward still evaluated (tictactoe_env.py:28-31), action >= 9 reports VALUE_ERROR.
__device__ __forceinline__ void ttt_step_fast(u64& own, u64& opp, int action, int player, int done_in, const u64* __restrict__ lines,
                                              int& reward, int& done_out, int& code, unsigned& valid) {
    const unsigned o = (unsigned)own, e = (unsigned)opp;             // 9-bit boards
    const bool in_range = (unsigned)action < 9u;
    const bool step = in_range && !done_in;
    code = action < 0 ? SPX_ENV_SKIPPED : (done_in ? SPX_ENV_GAME_OVER : (step ? SPX_ENV_OK : SPX_ENV_VALUE_ERROR));
    const int a = in_range ? action : 0;
    const unsigned occ = o | e;
    const unsigned bit = (step ? (1u << a) : 0u) & ~occ;
    const bool mine = player > 0;
    const unsigned o2 = o | (mine ? bit : 0u), e2 = e | (mine ? 0u : bit);
    const unsigned m = mine ? o2 : e2;
    const u64 L = lines[a];
    const unsigned l01 = (unsigned)L, l23 = (unsigned)(L >> 32);
    const unsigned m2 = m | (m << 16);
    const unsigned t01 = ~m2 & l01, t23 = ~m2 & l23;                 // a 16-bit field is 0 iff that line is complete
    const bool win = !(t01 & 0xFFFFu) || !(t01 >> 16) || !(t23 & 0xFFFFu) || !(t23 >> 16);
    reward = (step && win) ? 1 : 0;
    const unsigned occ2 = o2 | e2;
    valid = ~occ2 & 0x1FFu;
    done_out = step ? (int)((reward != 0) | (occ2 == 0x1FFu)) : done_in;
    own = (own & ~0x1FFULL) | o2;
    opp = (opp & ~0x1FFULL) | e2;
}

template <int GAME>
__device__ __forceinline__ void step_fast(u64& own, u64& opp, int action, int player, int done_in, const u64* __restrict__ lines,
                                          int& reward, int& done_out, int& code, unsigned& valid) {
    if (GAME == SPX_GAME_CONNECT4) c4_step_fast(own, opp, action, player, done_in, lines, reward, done_out, code, valid);
    else ttt_step_fast(own, opp, action, player, done_in, lines, reward, done_out, code, valid);
}

template <int GAME>
__host__ __device__ __forceinline__ unsigned valid_mask(u64 own, u64 opp) {
    typedef Rules<GAME> R;
    u64 occ = own | opp;
    if (GAME == SPX_GAME_CONNECT4) {
        unsigned m = 0;
#pragma unroll
        for (int c = 0; c < 7; ++c) m |= (unsigned)(((occ >> (7 * c + 5)) & 1ULL) ^ 1ULL) << c;  // heights < 6
        return m;
    }
    return (unsigned)(~occ & R::BOARD);  // board.reshape(-1) == 0
}

// One env.step.  `own`/`opp` are the +1 / -1 cells of the env's frame; `player` is +1 or -1.
// Returns SPX_ENV_OK / SPX_ENV_VALUE_ERROR; reward in {0,1}; done = reward != 0 or board full.
template <int GAME>
__host__ __device__ __forceinline__ int env_step(u64& own, u64& opp, int action, int player, int& reward, int& done) {
    typedef Rules<GAME> R;
    u64 occ = own | opp;
    int x, y;
    if ((unsigned)action >= (unsigned)R::A) { reward = 0; done = 0; return SPX_ENV_VALUE_ERROR; }  // the reference raises IndexError
    if (GAME == SPX_GAME_CONNECT4) {
        x = action;
        y = SPX_POPC((occ >> (7 * action)) & 0x3FULL);  // heights[action]
        if (y >= R::H) { reward = 0; done = 0; return SPX_ENV_VALUE_ERROR; }
        u64 bit = 1ULL << (x * 7 + y);
        if (player > 0) own |= bit; else opp |= bit;
    } else {
        x = action / 3;  // np.unravel_index(action, (3,3))
        y = action % 3;
        u64 bit = 1ULL << action;
        if (!(occ & bit)) { if (player > 0) own |= bit; else opp |= bit; }  // occupied: silent no-op
    }
    u64 mover = player > 0 ? own : opp;
    reward = reward_at<GAME>(mover, x, y);
    done = (reward != 0) || (SPX_POPC(own | opp) == R::CELLS);
    return SPX_ENV_OK;
}

}  // namespace spx

// spx_tower.cu -- the policy/value network evaluated at the MCTS leaves, hand-written for sm_100a.
//
// Replaces InferenceWorker.calculate (inference_worker.py:114-119) + ResidualTower.forward
// (games/general/modules.py:27-40,88-107): preprocess -> conv3x3 stem -> N residual blocks @128 ch ->
// 1x1 policy/value head convs (BN folded, eval mode) -> Linear+softmax / Linear-ReLU-Linear-tanh heads in ONE
// persistent kernel (tower_kernel<2>, the default).  DESIGN.md 3.3/3.4 has the measurements behind every choice.
//
// Design (B200-first):
//   * A cluster of two CTAs (one SM pair); every CTA owns 7 boards for the whole network.  Activations never leave
//     shared memory: two bf16 ping-pong buffers of 400 rows x 128 channels (row = padded board cell, index col*7+row,
//     56 rows per board incl. zero guard cells, so a 3x3 tap is a constant row shift and needs no im2col).
//   * Every conv is 9 (taps) x 8 (K slices of 16 channels) tcgen05.mma.cta_group::2 steps of M=256 (128 rows of each CTA)
//     x N=128 out-channels x K=16, for 3 row tiles; A = activations read in place through a shifted no-swizzle K-major
//     shared-memory descriptor, B = the weight slice of that (tap, K slice) -- each CTA stages half of it --, D = fp32
//     accumulators in TMEM (3 x 128 columns).  The folded-BN bias is the first MMA of a layer (ones x bias^T).
//   * Weights (13.5 MB bf16 for 20 blocks, L2 resident) are pre-packed on the host in exactly the order the MMAs consume
//     them and streamed by one producer lane per CTA with cp.async.bulk (TMA) through a 6 x 4 KB mbarrier ring.
//   * One elected lane of the leader CTA issues every MMA; its instruction stream paces the kernel, so descriptors are
//     advanced with single 32-bit adds.  Layers are handed over tile by tile (epi_done[t] / acc_full[t]): the next layer's
//     first tap starts on tile 0 while the epilogue still works on tiles 1 and 2.
//   * 16 epilogue warps read TMEM (tcgen05.ld 32x32b.x32, double-buffered), add the residual, apply ReLU + bf16 packing,
//     zero the padding rows and write the rows back into the other activation buffer.
//   * FC heads in the same kernel: the value layer on tcgen05 with the roles swapped (hidden units = M, boards = N) from a
//     second TMA ring in the dead activation buffer, the policy layer in fp32 on the epilogue warps, DSMEM exchange of the
//     two CTAs' half sums.
// Roles: warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warp 2 = peer relay (CTA 1), warps 4..19 = epilogue.
// Fallbacks: tower_kernel<1> (SPX_TOWER_NCTA=1: one CTA per 7 boards, cta_group::1) and heads_kernel
// (SPX_TOWER_FUSED_HEADS=0 or after tower_kernel<1>).  SPX_DBG_* macros build timing experiments only.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>

#include "spx_advance.cuh"

namespace spx {
int set_err(int code, const char* fmt, const char* detail);
void count_launch();
}  // namespace spx

#define SPX_CUDA_T(expr)                                                                  \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) return spx::set_err(SPX_E_CUDA, #expr ": %s", cudaGetErrorString(_e)); \
    } while (0)

namespace spx {
namespace tower {

// ------------------------------------------------------------------------------------------------ geometry
constexpr int CH = 128;          // trunk channels (filter_factor * 4)
constexpr int HEAD_CH = 64;      // 32 policy + 32 value head channels, one fused 1x1 conv
constexpr int BOARD_W = 7, BOARD_H = 6, CELLS = 42;
constexpr int PAD_STRIDE = 7;    // padded column stride (== the bitboard bit index)
constexpr int BOARD_ROWS = 56;   // rows per board incl. the shared zero guard column
constexpr int NB = 7;            // boards per CTA
constexpr int MT = 3;            // 128-row MMA tiles per CTA
constexpr int ROWS = MT * 128;   // 384 >= 6*56 + 48
constexpr int GUARD = 8;         // zero rows before/after (largest tap shift is +-8)
constexpr int ROWS_TOT = ROWS + 2 * GUARD;
constexpr int CHUNK_BYTES = ROWS_TOT * 16;   // one 8-channel chunk of every row
constexpr int ACT_BYTES = CHUNK_BYTES * (CH / 8);
#ifndef SPX_NSTAGE
#define SPX_NSTAGE 3
#endif
constexpr int NSTAGE = SPX_NSTAGE;
constexpr int KSTEP_BYTES = 2 * CH * 16;     // one K=16 weight slice: [2 k-chunks][128 out][8 in] bf16 = 4 KB
constexpr int STAGE_BYTES = 2 * KSTEP_BYTES; // a ring stage carries two K steps (K = 32)
#ifndef SPX_EPI_WARPS
#define SPX_EPI_WARPS 16
#endif
constexpr int EPI_WARP0 = 4, EPI_WARPS = SPX_EPI_WARPS, EPI_THREADS = EPI_WARPS * 32, EPI_SPLIT = EPI_WARPS / 4;
constexpr int NUM_THREADS = EPI_WARP0 * 32 + EPI_THREADS;
static_assert(EPI_SPLIT == 4, "epilogue column split is written for 4 parts of 32 (trunk) / 16 (head) columns");
constexpr int FLAT = 32 * CELLS;             // 1344 inputs of each head's first Linear
constexpr int FC_HIDDEN = 256;
constexpr int FC_KSTEPS = FLAT / 16;                                   // 84 K steps of the value layer Linear(1344 -> 256)
constexpr int FC_KSTEP_BYTES = 2 * (FC_HIDDEN / 2) * 16;               // per CTA half: [2 k-chunks][128 hidden][8] bf16 = 4 KB
constexpr size_t FC_STREAM_BYTES = (size_t)FC_KSTEPS * 2 * FC_KSTEP_BYTES;   // fused value layer, appended to the conv stream
__host__ __device__ constexpr size_t bias_slices_bytes(int num_blocks) { return (size_t)(2 * num_blocks + 1) * (2 * CH * 16) + (size_t)(2 * HEAD_CH * 16); }
// fused FC heads (SM-pair kernel): after the last trunk layer activation buffer 0 is dead and becomes a 12 x 8 KB weight
// ring for the value layer; after the head conv buffer 1 is dead and holds the head activations
constexpr int FC_STAGE_BYTES = 2 * FC_KSTEP_BYTES;                     // two K steps per stage (per CTA)
constexpr int FC_STAGES = 12;
constexpr int FC_ITERS = FC_KSTEPS / 2;                                // 42 ring stages
constexpr int XV_OFF = 0;                                              // value activations, bf16 [168 k-chunks][8 boards][8]
constexpr int XV_BYTES = (FLAT / 8) * 128;
constexpr int XP_OFF = 24576;                                          // policy activations, fp32 [7 boards][1344]
constexpr int XP_BYTES = NB * FLAT * 4;
constexpr int FCS_OFF = 65536;                                         // fp32 scratch: logits [7][8], partial [4][16], own [16], peer [16]
static_assert(FC_STAGES * FC_STAGE_BYTES <= ACT_BYTES && XV_OFF + XV_BYTES <= XP_OFF && XP_OFF + XP_BYTES <= FCS_OFF && FCS_OFF + 1024 <= ACT_BYTES, "fused-heads buffers");


// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// long waits (a whole layer of MMAs): back off between polls so that 500 idle threads do not compete for issue slots / power
#ifndef SPX_WAIT_SLEEP_NS
#define SPX_WAIT_SLEEP_NS 20   // measured in the loop: 0.440 -> 0.435 ms per forward (the run is power limited, idle polling costs clock)
#endif
__device__ __forceinline__ void mbar_wait_backoff(void* bar, unsigned parity) {
#if SPX_WAIT_SLEEP_NS > 0
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (!done) __nanosleep(SPX_WAIT_SLEEP_NS);
    }
#else
    mbar_wait(bar, parity);
#endif
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(unsigned taddr, unsigned (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32_nowait(unsigned taddr, unsigned (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_st32(unsigned taddr, const unsigned (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
          "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]),
          "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]),
          "r"(v[30]), "r"(v[31])
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// the folded-BN bias of a layer is pre-stored into the accumulator columns by the epilogue warps, so every MMA accumulates
__device__ __forceinline__ void store_bias_to_tmem(unsigned tcol, const float* bias_s) {
    unsigned b[32];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float4 f = *reinterpret_cast<const float4*>(bias_s + 4 * k);
        b[4 * k] = __float_as_uint(f.x); b[4 * k + 1] = __float_as_uint(f.y); b[4 * k + 2] = __float_as_uint(f.z); b[4 * k + 3] = __float_as_uint(f.w);
    }
#pragma unroll
    for (int t = 0; t < MT; ++t) tc_st32(tcol + (unsigned)(t * 128), b);
    tc_wait_st();
}
// max(x, 0) and round-to-nearest packing of two activations in one instruction: bf16, or fp16 (F16 = true: 11 instead of 8
// significant bits -- the reference's own GPU arithmetic is fp16 autocast, inference_worker.py:117; saturating, so that a
// value beyond 65504 stays finite)
template <bool F16> __device__ __forceinline__ unsigned relu_pack2(float lo, float hi) {
    unsigned r;
    if constexpr (F16) asm("cvt.rn.satfinite.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    else asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <bool F16> __device__ __forceinline__ float2 unpack2(unsigned w) {
    if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2*>(&w));
    else return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xFFFF0000u));
}
__device__ __forceinline__ void tc_ld16(unsigned taddr, unsigned (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor: core matrix = 8 rows x 16 bytes (contiguous 128 B);
// SBO = byte stride between 8-row groups, LBO = byte stride between the two 16-byte K chunks of one MMA.
__device__ __forceinline__ unsigned long long make_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
    return (unsigned long long)((saddr >> 4) & 0x3FFFu) | ((unsigned long long)((lbo >> 4) & 0x3FFFu) << 16) |
           ((unsigned long long)((sbo >> 4) & 0x3FFFu) << 32) | (1ULL << 46);
}
// instruction descriptor: D=f32 (bit 4), A and B formats at bits 7 / 10 (0 = fp16, 1 = bf16), both K-major, N>>3 at bit 17,
// M>>4 at bit 24
__host__ __device__ constexpr unsigned make_idesc(int M, int N, bool f16 = false) {
    return (1u << 4) | (f16 ? 0u : ((1u << 7) | (1u << 10))) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

// Which padded row is a real board cell, and which one.  Connect4 fills the 7x6 slot; a TicTacToe board (ResidualTower on
// the 3x3 env, main.py:74 with --g tictactoe) is embedded in the corner of the same slot: cell (x, y) sits at column x, row y,
// every other row of the slot is forced to zero by the epilogue like the padding rows, which is exactly the zero padding of
// a 3x3 convolution.  `cell` is the flatten index of modules.py:98,103 (x * H + y), `bit` the bitboard index of the cell.
__device__ __forceinline__ bool row_is_cell(int game, int row, int& board, int& cell, int& bit) {
    board = row / BOARD_ROWS;
    const int p = row - board * BOARD_ROWS;
    const int col = p / PAD_STRIDE, r = p - col * PAD_STRIDE;
    if (game == SPX_GAME_TICTACTOE) { cell = col * 3 + r; bit = cell; return col < 3 && r < 3; }
    cell = col * BOARD_H + r; bit = p;
    return p < BOARD_W * PAD_STRIDE && r < BOARD_H;
}

struct LayerInfo { int taps, kslices, n, in_buf, out_buf, residual; };
__device__ __forceinline__ LayerInfo layer_info(int l, int n_layers) {
    LayerInfo li;
    if (l == 0) { li.taps = 9; li.kslices = 1; li.n = CH; li.in_buf = 0; li.out_buf = 1; li.residual = 0; }
    else if (l == n_layers - 1) { li.taps = 1; li.kslices = CH / 16; li.n = HEAD_CH; li.in_buf = 1; li.out_buf = -1; li.residual = 0; }
    else { const int second = ((l - 1) & 1); li.taps = 9; li.kslices = CH / 16; li.n = CH; li.in_buf = second ? 0 : 1; li.out_buf = second ? 1 : 0; li.residual = second; }
    return li;
}
__device__ __forceinline__ int tap_shift(int tap) { return (tap / 3 - 1) * PAD_STRIDE + (tap % 3 - 1); }  // (kh-1)*7 + (kw-1)

// ------------------------------------------------------------------------------------------------ the tower kernel
// One template, two variants:
//   NCTA == 2 (default): two CTAs (one SM pair, cluster of 2) share every weight slice: tcgen05.mma.cta_group::2 with
//     M=256 (128 rows of each CTA) x N=128; each CTA stages only HALF of the slice (64 output channels) and the pair
//     exchanges B through the tensor-core datapath, which halves the weight bytes written into / read from shared memory
//     per CTA and doubles the ring depth (6 x 4 KB).  The leader CTA (cluster rank 0) issues all MMAs; completion is
//     multicast to both CTAs' barriers; the peer relays "my half has landed" and "my epilogue is done" with remote
//     mbarrier arrives.
//   NCTA == 1 (SPX_TOWER_NCTA=1, fallback): every CTA is on its own (cta_group::1, M=128, 3 x 8 KB ring).
#ifdef SPX_DBG_TRACE
__device__ long long g_trace[64 * 16 + 3 * 160];   // + per-CTA {entry, exit, SM id} for the first 160 CTAs
#define SPX_TRACE(l, k) do { if (blockIdx.x == 0 && (l) < 64 && trace_on) g_trace[(l) * 16 + (k)] = clock64(); } while (0)
#define SPX_TRACE_IF(c, l, k) do { if (c) SPX_TRACE(l, k); __syncwarp(); } while (0)
__device__ __forceinline__ unsigned long long globaltimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ long long g_tt[64 * 12];   // fused tick kernel: globaltimer (ns) of 12 points of passes 20..83, CTA 0, first epilogue thread
#define SPX_TT(k) do { if (ENGINE && blockIdx.x == 0 && tid == EPI_WARP0 * 32 && pass >= 20 && pass < 84) g_tt[(pass - 20) * 12 + (k)] = (long long)clock64(); } while (0)
#else
#define SPX_TRACE(l, k) do { } while (0)
#define SPX_TRACE_IF(c, l, k) do { } while (0)
#define SPX_TT(k) do { } while (0)
#endif
template <int NCTA> struct SmemT {
    static constexpr int STAGES = NCTA * NSTAGE;
    unsigned char act[2][ACT_BYTES];
    unsigned char wstage[STAGES][STAGE_BYTES / NCTA];
    unsigned long long full[STAGES], empty[STAGES], peer_full[STAGES], acc_full[MT], epi_done[MT];   // accumulator complete / tile handed over, per 128-row tile
    unsigned long long fc_full[FC_STAGES], fc_empty[FC_STAGES], fc_peer_full[FC_STAGES], fc_done;   // fused FC heads
    unsigned long long own[NB], opp[NB];
    unsigned long long leaf2[8][2];     // ENGINE, fast mode: {own, opp} of slot w's next leaf, written by engine_step
    // fused tick kernel (ENGINE): leaves / hand-over state of the CTA's games, see the ENGINE comment above tower_kernel
    unsigned long long xbar;            // "the peer CTA's halves of my boards' hidden-layer sums have landed" (st.async complete_tx)
    float xval[8];
    int slot_status[8];                 // FS_* per board slot (fast mode)
    int quit, passes_done, searches_done[2];   // searches_done[k]: jobs finished by shadow warp k (shadow-all mode)
    int go;                             // ENGINE: ticks known to exist | GO_FINAL (no more after those), see "work-conserving launches"
    unsigned char need[8];
    alignas(16) float bias[NCTA == 1 ? 3 : 1][NCTA == 1 ? CH : 4];   // single-CTA kernel only: staged fp32 bias, read as float4
    alignas(128) unsigned char ones[256];   // SM-pair kernel: the A operand of the bias MMA (see issue_bias)
    unsigned tmem_base;
};

__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the same barrier in CTA `cta` of the cluster.  Waits stay plain CTA-scope try_wait: an acquire.cluster wait
// on these barriers cost +55 % kernel time.
__device__ __forceinline__ void mbar_arrive_remote(void* bar, unsigned cta) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\t"
                 "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}
template <int NCTA> __device__ __forceinline__ void tc_commit_t(void* bar) {
    if constexpr (NCTA == 2)   // completion of this thread's cta_group::2 MMAs -> the same barrier in both CTAs
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                     "h"((unsigned short)3) : "memory");
    else
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] += A[smem] * B[smem]^T, bf16 inputs, fp32 accumulate (the accumulators start from the pre-stored bias)
template <int NCTA> __device__ __forceinline__ void tc_mma_t(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned idesc) {
    if constexpr (NCTA == 2)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                     "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
                     "l"(adesc), "l"(bdesc), "r"(idesc) : "memory");
}

constexpr unsigned DESC_HI = (128u >> 4) | (1u << 14);                  // SBO = 128 B (bits 32..45), descriptor version 1 (bit 46)
constexpr unsigned A_DESC_FIELDS = (unsigned)(CHUNK_BYTES >> 4) << 16;   // LBO of the activation operand = chunk stride
// same MMA, descriptors passed as (low word, shared high word)
template <int NCTA> __device__ __forceinline__ void tc_mma_lo(unsigned d_tmem, unsigned a_lo, unsigned b_lo, unsigned idesc) {
    if constexpr (NCTA == 2)
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                     "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(idesc) : "memory");
    else
        asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, 1, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(idesc) : "memory");
}

// cta_group::2 MMA with a run-time accumulate flag (the value layer starts its accumulators from zero)
__device__ __forceinline__ void tc_mma_lo_acc2(unsigned d_tmem, unsigned a_lo, unsigned b_lo, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(DESC_HI), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void st_shared_remote_f32(void* local_addr, unsigned cta, float v) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tst.shared::cluster.f32 [ra], %2;\n\t}" ::"r"(smem_u32(local_addr)), "r"(cta), "f"(v) : "memory");
}

// one fp32 into the peer CTA's shared memory, completing 4 bytes of the transaction count of the peer's mbarrier `bar`
__device__ __forceinline__ void st_async_remote_f32(void* local_addr, void* bar, unsigned cta, float v) {
    asm volatile("{\n\t.reg .b32 ra, rb;\n\tmapa.shared::cluster.u32 ra, %0, %2;\n\tmapa.shared::cluster.u32 rb, %1, %2;\n\t"
                 "st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [ra], %3, [rb];\n\t}" ::"r"(smem_u32(local_addr)), "r"(smem_u32(bar)),
                 "r"(cta), "r"(__float_as_uint(v)) : "memory");
}
__device__ __forceinline__ void st_shared_remote_s32(void* local_addr, unsigned cta, int v) {
    asm volatile("{\n\t.reg .b32 ra;\n\tmapa.shared::cluster.u32 ra, %0, %1;\n\tst.volatile.shared::cluster.s32 [ra], %2;\n\t}" ::"r"(smem_u32(local_addr)), "r"(cta), "r"(v) : "memory");
}
constexpr int GO_FINAL = 0x40000000;
__device__ __forceinline__ int ld_volatile_s32(const int* p) { return *reinterpret_cast<const volatile int*>(p); }
__device__ __forceinline__ void st_volatile_s32(int* p, int v) { *reinterpret_cast<volatile int*>(p) = v; }
// warp-uniform read of a flag another warp may be writing right now (lanes must not disagree: the callers shuffle afterwards)
__device__ __forceinline__ int ld_flag_uniform(const int* p) { return __shfl_sync(0xffffffffu, ld_volatile_s32(p), 0); }

enum { FS_FAST = 0, FS_TODO = 1, FS_DONE = 2 };   // who works on a game slot: its epilogue warp / the shadow warp (pending, finished)

// SM-pair kernel: the folded-BN bias enters the accumulators through the tensor pipe instead of a tcgen05.st by the epilogue
// warps: D = ones * bias^T with accumulate = 0 is the first MMA of every layer.  A = `ones`: ONE 8-row core matrix whose
// rows are (1, 1, 0, ..., 0) aliased by all 16 row groups (SBO = 0) + a zero core matrix for the second k-chunk; B = a
// [2 k-chunks][n][8] slice of the weight stream with k = 0 -> bf16(bias), k = 1 -> bf16(bias - bf16(bias)) (~16 mantissa bits).
__device__ __forceinline__ void tc_mma_bias2(unsigned d_tmem, unsigned a_lo, unsigned b_lo, unsigned idesc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, 0, 0;\n\tmov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %4};\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(b_lo), "r"(1u << 14), "r"(DESC_HI), "r"(idesc) : "memory");
}

// All MMAs of one conv layer: taps x KPAIRS ring stages of KSTEPS K-steps x 3 row tiles.
template <int NCTA, int KSTEPS, int KPAIRS>
__device__ __forceinline__ void issue_layer(SmemT<NCTA>& S, int taps, unsigned a_lo_layer, unsigned b_fields, unsigned kstep16, unsigned idesc,
                                            unsigned tmem_base, bool leader, unsigned& stage, unsigned& sphase, int first_tap = 0) {
    constexpr int STAGES = SmemT<NCTA>::STAGES;
    for (int tap = first_tap; tap < taps; ++tap) {
#ifdef SPX_DBG_NO_SHIFT
        unsigned a_lo = a_lo_layer + 0u * (unsigned)tap;
#else
        unsigned a_lo = a_lo_layer + (unsigned)(taps == 9 ? tap_shift(tap) : 0);   // a row is 16 B = one descriptor address unit
#endif
#pragma unroll
        for (int kp = 0; kp < KPAIRS; ++kp) {
#if !defined(SPX_DBG_NO_TMA) && !defined(SPX_DBG_NO_FULL_WAIT)
            mbar_wait(&S.full[stage], sphase);
            if constexpr (NCTA == 2) mbar_wait(&S.peer_full[stage], sphase);
#endif
            const unsigned b_lo = b_fields | (smem_u32(S.wstage[stage]) >> 4);
            if (leader) {
                if (tap == taps - 1 && kp == KPAIRS - 1) {
                    // the layer's last stage, tile by tile: accumulator t is complete (acc_full[t]) 2 * KSTEPS * (2 - t) MMAs before
                    // the last one, so the epilogue starts on tile 0 while the tensor pipe finishes tiles 1 and 2
#pragma unroll
                    for (int t = 0; t < MT; ++t) {
#pragma unroll
                        for (int j = 0; j < KSTEPS; ++j)
                            tc_mma_lo<NCTA>(tmem_base + (unsigned)(t * 128), a_lo + (unsigned)(j * 2 * (CHUNK_BYTES >> 4) + t * 128),
                                            b_lo + (unsigned)j * kstep16, idesc);
                        tc_commit_t<NCTA>(&S.acc_full[t]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < KSTEPS; ++j)
#pragma unroll
                        for (int t = 0; t < MT; ++t)
                            tc_mma_lo<NCTA>(tmem_base + (unsigned)(t * 128), a_lo + (unsigned)(j * 2 * (CHUNK_BYTES >> 4) + t * 128),
                                            b_lo + (unsigned)j * kstep16, idesc);
                }
#ifndef SPX_DBG_NO_TMA
                tc_commit_t<NCTA>(&S.empty[stage]);   // frees the weight slot (in both CTAs) once these MMAs retire
#endif
            }
            __syncwarp();
            a_lo += (unsigned)(KSTEPS * 2 * (CHUNK_BYTES >> 4));
            if (++stage == STAGES) { stage = 0; sphase ^= 1u; }
        }
    }
}

// SM-pair kernel, layers with 128 input channels: bias MMAs + the FIRST tap (4 ring stages x 2 K steps x 3 tiles), issued per
// tile as the previous layer's epilogue releases the tiles one after the other (epi_done[t]).  The first tap shifts by -8
// rows (or 0 for the 1x1 head conv): tile t reads rows of tiles t-1 and t only, so tile 0 may start while tiles 1 and 2 are
// still in the epilogue.  MMAs of a tile that has just become available are interleaved with the MMAs the earlier tile has
// left (back-to-back MMAs into one accumulator run at ~150 instead of 64 cycles).  The 5 ring slots (bias slice + 4 stages)
// stay allocated until tile 2 has used them; the remaining taps run stage-major (issue_layer, first_tap = 1).
__device__ __forceinline__ void issue_first_tap_skewed(SmemT<2>& S, int taps, unsigned a_lo_layer, unsigned b_fields, unsigned kstep16, unsigned idesc,
                                                       unsigned tmem_base, bool leader, unsigned& stage, unsigned& sphase, unsigned ephase) {
    constexpr int STAGES = SmemT<2>::STAGES;
    constexpr unsigned KSTEP_A = 2 * (CHUNK_BYTES >> 4);
    const unsigned a_lo = a_lo_layer + (unsigned)(taps == 9 ? tap_shift(0) : 0);
    const unsigned ones_lo = ((128u >> 4) << 16) | (smem_u32(S.ones) >> 4);
    unsigned slot[5], b_lo[5];
    int have = 0;
    auto need = [&](int upto) {      // wait (whole warp) until ring stages 0..upto of this layer have landed in both CTAs
        for (; have <= upto; ++have) {
#ifndef SPX_DBG_NO_TMA
            mbar_wait(&S.full[stage], sphase);
            mbar_wait(&S.peer_full[stage], sphase);
#endif
            slot[have] = stage;
            b_lo[have] = b_fields | (smem_u32(S.wstage[stage]) >> 4);
            if (++stage == STAGES) { stage = 0; sphase ^= 1u; }
        }
    };
    auto bias = [&](int t) { tc_mma_bias2(tmem_base + (unsigned)(t * 128), ones_lo, b_lo[0], idesc); };
    auto mma = [&](int t, int s, int j) {
        tc_mma_lo<2>(tmem_base + (unsigned)(t * 128), a_lo + (unsigned)((s * 2 + j) * KSTEP_A + t * 128), b_lo[1 + s] + (unsigned)j * kstep16, idesc);
    };
    auto release = [&](int i) {
#ifndef SPX_DBG_NO_TMA
        tc_commit_t<2>(&S.empty[slot[i]]);
#endif
    };
    mbar_wait(&S.epi_done[0], ephase);
    tc_fence_after();
    need(2);
    if (leader) { bias(0); mma(0, 0, 0); mma(0, 0, 1); mma(0, 1, 0); mma(0, 1, 1); }
    __syncwarp();
    mbar_wait(&S.epi_done[1], ephase);
    tc_fence_after();
    need(4);
    if (leader) {
        bias(1);
        mma(1, 0, 0); mma(0, 2, 0); mma(1, 0, 1); mma(0, 2, 1);
        mma(1, 1, 0); mma(0, 3, 0); mma(1, 1, 1); mma(0, 3, 1);
    }
    __syncwarp();
    mbar_wait(&S.epi_done[2], ephase);
    tc_fence_after();
    if (leader) {
        bias(2); release(0);
        mma(2, 0, 0); mma(1, 2, 0); mma(2, 0, 1); mma(1, 2, 1); release(1);
        mma(2, 1, 0); mma(1, 3, 0); mma(2, 1, 1); mma(1, 3, 1); release(2);
        mma(2, 2, 0); mma(2, 2, 1); release(3);
        mma(2, 3, 0); mma(2, 3, 1); release(4);
    }
    __syncwarp();
}

// The fused tick kernel calls the search engine through this NON-INLINED function: advance_game gets the whole 96-register
// budget for its select loop (inlined into the epilogue code it competed with ~25 kernel-level values and ptxas spilled 60
// registers across it); the caller saves what it keeps live around the call, once per simulation.  The leaf goes straight to
// the two shared-memory words the network phase reads.
// SOFT: the instance the shadow warp calls -- it searches while the tensor pipe runs, and FP64 instructions next to tcgen05 MMAs
// slow them down, so its arithmetic runs on the integer pipe (spx_softf64.cuh)
#ifndef SPX_FAST_SOFT
#define SPX_FAST_SOFT false    // the epilogue warps' searches between two passes (tensor pipe idle): FP64 instructions (-DSPX_FAST_SOFT=true: integer pipe)
#endif
#ifndef SPX_SHADOW_SOFT
#define SPX_SHADOW_SOFT true   // -DSPX_SHADOW_SOFT=false: the shadow warp on the FP64 pipe (timing experiments)
#endif
template <int GAME, bool SOFT>
__device__ __noinline__ int engine_step(const spx::EngineDev& E, const int g, const float my_p, const float v_in, const int budget,
                                        const int defer_leaf, unsigned long long* leaf2) {
    const int lane = threadIdx.x & 31;
    const spx::AdvPre pre = spx::advance_prefetch<GAME>(E, g, lane);
    unsigned long long own, opp;
    const int flags = spx::advance_game<GAME, false, true, SOFT>(E, g, lane, pre, my_p, v_in, budget, defer_leaf != 0, false, own, opp);
    if (leaf2 && lane == 0) { leaf2[0] = own; leaf2[1] = opp; }
    return flags;
}

// GAME is a template parameter: with the head sizes as run-time values the Connect4 kernel spilled (712-byte stack frame at the
// 96-register cap, +13 % instructions, 8 MB of local-memory write-back per launch in ncu)
// ENGINE = true is the fused tick kernel (spx_tick_fused): the kernel runs `n_ticks` ticks without the host; before every network
// evaluation of a board group the first NB epilogue warps of the CTA that owns the group run the search engine's per-game state
// machine (advance_game, spx_advance.cuh) for the group's games, i.e. consume the previous outputs and produce the next leaves.
// Games are independent, so clusters never wait for each other: no grid-wide barrier, no launch gaps, no advance-kernel tail.
template <int NCTA, int GAME = SPX_GAME_CONNECT4, bool ENGINE = false, bool F16 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1)
tower_kernel(const unsigned long long* __restrict__ own_g, const unsigned long long* __restrict__ opp_g,
             const unsigned char* __restrict__ needs, long long n_boards, int n_layers,
             const unsigned char* __restrict__ wconv, const float* __restrict__ bias_all, float* __restrict__ head_out,
             int fused_in, const float* __restrict__ polw, const float* __restrict__ polb, const float* __restrict__ fc_b1,
             const float* __restrict__ fc_w2, const float* __restrict__ fc_b2, float* __restrict__ policy_out, float* __restrict__ value_out,
             const __grid_constant__ EngineDev E, const int n_ticks, const long long pass_budget) {
    typedef SmemT<NCTA> Smem;
    // leaves / needs_eval are produced inside this launch when ENGINE: read them through L2 (__ldcg), never through the
    // read-only path the __restrict__ const parameters allow
    auto ld_need = [&](long long gb) -> bool { return needs == nullptr || needs[gb]; };   // plain forward (ENGINE: S.need)
    const bool fused = NCTA == 2 && fused_in != 0;   // FC heads inside this kernel (no head_out round trip, no second launch)
    constexpr int game = GAME;
    constexpr int cells = GAME == SPX_GAME_TICTACTOE ? 9 : CELLS, n_act = GAME == SPX_GAME_TICTACTOE ? 9 : 7;
    constexpr int flat = 32 * cells, fc_iters = flat / 32;   // inputs of each head's first Linear; ring stages of the value layer
    constexpr int STAGES = Smem::STAGES;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    Smem& S = *reinterpret_cast<Smem*>(smem_raw);
    const unsigned crank = NCTA == 2 ? cluster_ctarank() : 0u;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long n_groups = (n_boards + NB - 1) / NB;
    const bool trace_on = true;   // (-DSPX_DBG_TRACE) the fused tick kernel traces the layers of ONE mid-launch pass, see the unit loop
    (void)trace_on;
    if (tid == 0) SPX_TRACE(62, 0);
#ifdef SPX_DBG_TRACE
    if (tid == 0 && blockIdx.x < 160) { unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); g_trace[1024 + 3 * blockIdx.x] = (long long)globaltimer_ns(); g_trace[1024 + 3 * blockIdx.x + 2] = smid; }
#endif

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&S.full[s], 1); mbar_init(&S.empty[s], 1); mbar_init(&S.peer_full[s], 1); }
        for (int t = 0; t < MT; ++t) mbar_init(&S.acc_full[t], 1);
        for (int t = 0; t < MT; ++t) mbar_init(&S.epi_done[t], EPI_WARPS * NCTA);   // one arrive per epilogue warp of every CTA of the cluster
        for (int s = 0; s < FC_STAGES; ++s) { mbar_init(&S.fc_full[s], 1); mbar_init(&S.fc_empty[s], 1); mbar_init(&S.fc_peer_full[s], 1); }
        mbar_init(&S.fc_done, 1);
        mbar_init(&S.xbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if constexpr (NCTA == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_base)) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    // zero both activation buffers once: guard rows and padding cells must read as zero forever
    for (int i = tid; i < 2 * ACT_BYTES / 16; i += NUM_THREADS) reinterpret_cast<uint4*>(S.act[0])[i] = make_uint4(0, 0, 0, 0);
    if (tid < 16) reinterpret_cast<uint4*>(S.ones)[tid] = tid < 8 ? make_uint4(F16 ? 0x3C003C00u : 0x3F803F80u, 0, 0, 0) : make_uint4(0, 0, 0, 0);   // (1, 1, 0 x 6) x 8 rows | zeros
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    if constexpr (NCTA == 2) cluster_sync_all();   // both CTAs' barriers initialised and buffers zeroed before any remote arrive / MMA
    tc_fence_after();
    const unsigned tmem_base = S.tmem_base;
    if (tid == 0) SPX_TRACE(62, 1);

    unsigned stage = 0, sphase = 0;   // weight ring position (producer, relay and MMA issuer walk the same sequence)
    unsigned lphase = 0;              // per-layer parity of acc_full (n_layers, an even number, of completions per unit)
    unsigned ephase = 0;              // parity of epi_done (one more completion per unit when the FC heads are fused)
    unsigned fstage = 0, fphase = 0;  // FC weight ring position
    unsigned dphase = 0;              // fc_done parity (one completion per unit)
    bool first_unit = true;

    const long long n_units = (n_groups + NCTA - 1) / NCTA;      // a unit = the NCTA board groups one cluster works on together
    // ---- ENGINE (fused tick kernel): who advances the games.
    // A pass = one network evaluation of one unit.  The search of a game cannot overlap the pass that evaluates its own leaf,
    // but it needs nothing else, so:
    //  * a cluster that owns ONE unit (<= 1036 games per GPU) is in "fast mode": after every pass epilogue warp w (w < 7) runs ONE
    //    simulation attempt of game w (advance_game, budget 1: consume the outputs it has just computed, select, expand) -- a few
    //    microseconds between two passes.  Whatever does not end in a leaf (chains of terminal re-visits, which need no network)
    //    is handed to the SHADOW WARP (warp 3), which keeps simulating during the next pass and hands the game back with its leaf.
    //  * a cluster that owns SEVERAL units is in "shadow-all mode": the shadow warp advances the 7 games of unit u while the
    //    tensor pipe evaluates the other units (job p needs the outputs of pass p - U only), the epilogue warps never search.
    // All hand-overs go through shared-memory flags + fences inside the CTA; the two CTAs of a pair only exchange the hidden
    // layer's half sums (st.async + mbarrier).  No CTA-wide or cluster-wide barrier inside the tick loop.
    const long long ustride = gridDim.x / NCTA, unit_first = blockIdx.x / NCTA;
    const int U = ENGINE ? (unit_first < n_units ? (int)((n_units - 1 - unit_first) / ustride + 1) : 0) : 0;
    const bool shadow_all = ENGINE && U >= 2;
    float my_p = 0.f, v_next = 0.f;      // epilogue warps w < NB: the outputs of the pass for game w (this lane's prior, the value)
    if constexpr (ENGINE) {
        if (tid < 8) { S.slot_status[tid] = FS_FAST; S.need[tid] = 0; }
        // work-conserving launches (pass_budget >= 0, spx_tick_fused_balanced): tick 0 exists for every cluster; whether tick t + 1
        // exists is decided at the start of tick t by a ticket drawn from the launch's budget of passes (E.ticks[1]), see below
        if (tid == 0) { S.quit = 0; S.passes_done = 0; S.searches_done[0] = 0; S.searches_done[1] = 0; S.go = U == 0 ? GO_FINAL : pass_budget >= 0 ? 1 : (n_ticks | GO_FINAL); }
        __syncthreads();
        // Shadow-all mode has a second shadow warp: the warp of this CTA that has no role in the tick loop (the leader's relay warp,
        // the peer's issuer warp) takes games 4..6 of every job, warp 3 games 0..3 -- with the evaluation cache a unit's seven
        // searches in a row only just fitted under one pass of the other unit (two units per SM pair).
        const int sw = warp == 3 ? 0 : ((shadow_all && NCTA == 2 && warp == (crank == 0 ? 2 : 1)) ? 1 : -1);
        if (sw >= 0) {
            // ===================== shadow warp(s): simulations that need no network evaluation of the running pass
            const int j0 = (shadow_all && sw == 1) ? 4 : 0, j1 = (shadow_all && sw == 0) ? 4 : NB;
            int p = 0, j = j0;
            for (;;) {
                long long gb = 0;
                bool run = false, defer = false;
                int budget = 0;
                float sp = 0.f, sv = 0.f;
                if (shadow_all) {
                    if (j == j0 && p % U == 0) {   // first job of tick p / U: does that tick exist?
                        int w;
                        for (;;) { w = ld_flag_uniform(&S.go); if (p / U < (w & (GO_FINAL - 1)) || (w & GO_FINAL)) break; __nanosleep(200); }
                        if (p / U >= (w & (GO_FINAL - 1))) break;
                    }
                    if (j == j0) {   // job p = the leaves of pass p: needs the outputs of the same unit's previous pass (p - U)
                        while (ld_flag_uniform(&S.passes_done) < p - U + 1) __nanosleep(200);
                        __threadfence();
                    }
                    gb = (NCTA * (unit_first + (long long)(p % U) * ustride) + crank) * NB + j;
                    run = gb < n_boards;
                    budget = 2 * E.cfg.max_sims_per_tick;
                    if (run) {
                        sp = lane < n_act ? __ldcg(policy_out + gb * n_act + lane) : 0.f;
                        sv = __ldcg(value_out + gb);
                    }
                } else {
                    if (j == 0 && ld_flag_uniform(&S.quit)) break;
                    gb = (NCTA * unit_first + crank) * NB + j;
                    run = ld_flag_uniform(&S.slot_status[j]) == FS_TODO;
                    if (run) __threadfence();
                    budget = 4 * E.cfg.max_sims_per_tick;
                    defer = true;       // the running pass does not evaluate this leaf: it waits for the next one
                }
                if (run && shadow_all && (E.cfg.reserved0 & 1)) run = false;   // timing experiment (SPX_DBG_FLAGS=1): no search, the same leaves again
#ifdef SPX_DBG_TRACE
                if (shadow_all && (E.cfg.reserved0 & 1) && (E.cfg.reserved0 & 0xFF8) && gb < n_boards) {
                    // synthetic interference instead of the search (about 12 us per game slot): which resource does a busy warp take
                    // from the tensor pipe?  8: dependent fp64 math, 16: dependent integer math, 32: dependent L2 loads
                    const long long t_end = clock64() + 20000;
                    double xd = 1.0 + lane;
                    unsigned xi = lane + 1;
                    const char* pool = E.pool + (size_t)gb * 2 * (size_t)E.nodes_per_tree * 160;
                    unsigned off = lane * 32;
                    while (clock64() < t_end) {
                        if (E.cfg.reserved0 & 8) { for (int i = 0; i < 32; ++i) xd = __ddiv_rn(__dadd_rn(xd, 1.5), __dsqrt_rn(__dadd_rn(xd, 2.0))); }
                        if (E.cfg.reserved0 & 64) { for (int i = 0; i < 64; ++i) xd = __dadd_rn(xd, 1.5); }
                        if (E.cfg.reserved0 & 128) { for (int i = 0; i < 64; ++i) xd = __dmul_rn(xd, 1.0000001); }
                        if (E.cfg.reserved0 & 256) { for (int i = 0; i < 64; ++i) xd = __fma_rn(xd, 1.0000001, 0.25); }
                        if (E.cfg.reserved0 & 512) { for (int i = 0; i < 16; ++i) xd = __dadd_rn(__dmul_rn(xd, 1.0000001), (double)(float)xi); for (int i = 0; i < 200; ++i) xi = xi * 1664525u + 1013904223u; }
                        if (E.cfg.reserved0 & 1024) {   // 64 KB of straight-line integer code: instruction-cache footprint
#pragma unroll 4096
                            for (int i = 0; i < 4096; ++i) xi = xi * (1664525u + 2u * (unsigned)i) + 1013904223u;
                        }
                        if (E.cfg.reserved0 & 2048) {   // 64-bit integer multiplies, shuffles, votes: what the integer-pipe arithmetic is made of
                            u64 z = xi | ((u64)xi << 32);
                            for (int i = 0; i < 64; ++i) { z = __umul64hi(z | 1ULL, 0x9E3779B97F4A7C15ULL) + z * 3ULL; z ^= __shfl_xor_sync(0xffffffffu, z, 1); z += __clzll((long long)z); }
                            xi = (unsigned)z ^ (unsigned)__ballot_sync(0xffffffffu, z & 1);
                        }
                        if (E.cfg.reserved0 & 16) { for (int i = 0; i < 256; ++i) xi = xi * 1664525u + 1013904223u; }
                        if (E.cfg.reserved0 & 32) { for (int i = 0; i < 8; ++i) off = (off * 1664525u + 1013904223u + (unsigned)__ldcg((const int*)(pool + (off & 0xFFFE0)))) ; }
                    }
                    if (xd == 0.123 || xi == 77u || off == 0x12345u) E.leaf_own[gb] = 1;   // keep the loops alive
                }
#endif
                if (run && shadow_all && (E.cfg.reserved0 & 2) && crank == 0) run = false;   // timing experiment (SPX_DBG_FLAGS=2): only the peer CTA searches
                if (run && shadow_all && (E.cfg.reserved0 & 4) && crank == 1) run = false;   // timing experiment (SPX_DBG_FLAGS=4): only the leader CTA searches
                if (run) {
                    const int flags = engine_step<GAME, SPX_SHADOW_SOFT>(E, (int)gb, sp, sv, budget, defer ? 1 : 0, nullptr);
                    if (!shadow_all && flags) {   // a leaf, or nothing more to do: back to the epilogue warp (0 = budget used up: go on later)
                        __threadfence();
                        if (lane == 0) st_volatile_s32(&S.slot_status[j], FS_DONE);
                    }
                }
                __syncwarp();
#ifdef SPX_DBG_SHADOW_SLEEP   // timing experiment: the shadow warp's work spread out in time
                __nanosleep(SPX_DBG_SHADOW_SLEEP);
#endif
                if (++j == j1) {
                    j = j0;
                    if (shadow_all) { __threadfence(); if (lane == 0) st_volatile_s32(&S.searches_done[sw], p + 1); ++p; }
                    else __nanosleep(300);
                }
            }
        }
    }
    // warps without a role in the tick loop (and the shadow warp, whose loop is above) go straight to the final barrier
    const bool has_role = warp == 0 || (warp == 1 && crank == 0) || (NCTA == 2 && warp == 2 && crank == 1) || warp >= EPI_WARP0;
    bool go_final = false;               // decider thread: the last tick of this cluster is known
    for (int tick = 0; ENGINE ? has_role : tick < 1; ++tick) {
    if constexpr (ENGINE) {   // does tick `tick` exist?  (exact launches: S.go = n_ticks | GO_FINAL from the start)
        int w;
        for (;;) { w = ld_flag_uniform(&S.go); if (tick < (w & (GO_FINAL - 1)) || (w & GO_FINAL)) break; __nanosleep(100); }
        if (tick >= (w & (GO_FINAL - 1))) break;
    }
    int ui = 0;
    for (long long unit = blockIdx.x / NCTA; unit < n_units; unit += gridDim.x / NCTA, ++ui) {
        const long long grp = NCTA * unit + crank;
        const int pass = tick * U + ui;
        (void)pass;
#ifdef SPX_DBG_TRACE
        const bool trace_on = !ENGINE || pass == 30;   // a pass in the middle of the launch: the shadow warp is searching next to it
#endif
        // cluster-uniform skip when none of the boards of this unit asked for an evaluation (plain forward only: the fused tick
        // kernel always evaluates -- agreeing on a skip would cost a cluster round trip per tick)
        bool any = ENGINE || needs == nullptr;
        if (!any) for (int b = 0; b < NCTA * NB; ++b) { long long gb = unit * NCTA * NB + b; if (gb < n_boards && ld_need(gb)) any = true; }
        if (!any) continue;

        if (warp == 0) {
            // ===================== TMA producer: stream the pre-packed weight slices in consumption order.
            // The whole warp walks the loop (warp-uniform control flow keeps addresses in uniform registers);
            // one elected lane issues the copies.
            const bool leader = elect_one();
            const unsigned char* wp = wconv;
            for (int l = 0; l < n_layers; ++l) {
                const LayerInfo li = layer_info(l, n_layers);
                const int ksteps = li.kslices >= 2 ? 2 : 1;                                   // K steps (of 16 channels) per ring stage
                const unsigned bytes = 2u * (unsigned)li.n * 16u * (unsigned)ksteps / NCTA;   // this CTA's share of the stage
                const int iters = li.taps * (li.kslices / ksteps);
                const unsigned bbytes = 2u * (unsigned)(li.n / 2) * 16u;     // one CTA's half of the layer's bias slice
                if constexpr (NCTA == 1) wp += 2 * bbytes;                    // the single-CTA kernel seeds the bias with tcgen05.st
                for (int it = (NCTA == 2 ? -1 : 0); it < iters; ++it) {
#ifdef SPX_DBG_NO_TMA
                    continue;
#endif
                    mbar_wait(&S.empty[stage], sphase ^ 1u);
                    if (it < 0) {   // the bias slice is the first ring stage of every layer
                        if (leader) {
                            mbar_expect_tx(&S.full[stage], bbytes);
                            tma_bulk_g2s(S.wstage[stage], wp + (size_t)crank * bbytes, bbytes, &S.full[stage]);
                        }
                        __syncwarp();
                        wp += 2 * bbytes;
                        if (++stage == STAGES) { stage = 0; sphase ^= 1u; }
                        continue;
                    }
                    if (leader) {
#ifdef SPX_DBG_FAKE_TMA
                        mbar_arrive(&S.full[stage]);        // timing experiment: the barrier protocol without the copy
#else
                        mbar_expect_tx(&S.full[stage], bytes);
                        tma_bulk_g2s(S.wstage[stage], wp + (size_t)crank * bytes, bytes, &S.full[stage]);
#endif
                    }
                    __syncwarp();
                    wp += (unsigned)NCTA * bytes;
                    if (++stage == STAGES) { stage = 0; sphase ^= 1u; }
                }
            }
            if (fused) {
                // value-layer weights into the dead activation buffer 0: its last readers are the MMAs of the last trunk layer
                // (layer n_layers-2, an even index: n_layers is even, so that completion of acc_full always has parity 0; this
                // warp is at most a ring (6 stages) ahead of the issuer, so the barrier cannot be a whole phase behind)
                mbar_wait(&S.acc_full[MT - 1], 0u);
                for (int it = 0; it < fc_iters; ++it) {
                    mbar_wait(&S.fc_empty[fstage], fphase ^ 1u);
                    if (leader) {
                        mbar_expect_tx(&S.fc_full[fstage], FC_STAGE_BYTES);
                        tma_bulk_g2s(S.act[0] + fstage * FC_STAGE_BYTES, wp + (size_t)crank * FC_STAGE_BYTES, FC_STAGE_BYTES, &S.fc_full[fstage]);
                    }
                    __syncwarp();
                    wp += 2 * FC_STAGE_BYTES;
                    if (++fstage == FC_STAGES) { fstage = 0; fphase ^= 1u; }
                }
            }
        } else if (NCTA == 2 && warp == 2 && crank == 1) {
            // ===================== peer relay: tell the leader when this CTA's half of each weight stage has landed
            const bool leader = elect_one();
            for (int l = 0; l < n_layers; ++l) {
                const LayerInfo li = layer_info(l, n_layers);
                const int ksteps = li.kslices >= 2 ? 2 : 1;
                const int iters = li.taps * (li.kslices / ksteps) + 1;   // + the bias slice
                for (int it = 0; it < iters; ++it) {
#ifdef SPX_DBG_NO_TMA
                    continue;
#endif
                    mbar_wait(&S.full[stage], sphase);
                    if (leader) mbar_arrive_remote(&S.peer_full[stage], 0);
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; sphase ^= 1u; }
                }
            }
            if (fused) {
                for (int it = 0; it < fc_iters; ++it) {
                    mbar_wait(&S.fc_full[fstage], fphase);
                    if (leader) mbar_arrive_remote(&S.fc_peer_full[fstage], 0);
                    __syncwarp();
                    if (++fstage == FC_STAGES) { fstage = 0; fphase ^= 1u; }
                }
            }
        } else if (warp == 1 && crank == 0) {
            // ===================== MMA issuer (leader CTA): warp-uniform loop, one elected lane issues tcgen05.mma / commit.
            // This single instruction stream paces the whole kernel (6 MMAs of 64 cycles per ring stage), so descriptors
            // are advanced with one 32-bit add each: the high word (SBO, version) is a constant, the low word is
            // LBO << 16 | address >> 4, and every operand of a stage is the stage base plus a compile-time constant.
            const bool leader = elect_one();
            for (int l = 0; l < n_layers; ++l) {
                const LayerInfo li = layer_info(l, n_layers);
                const unsigned idesc = make_idesc(128 * NCTA, li.n, F16);
                const unsigned b_lbo16 = (unsigned)li.n / NCTA;                       // (n/NCTA * 16 B) >> 4: stride between the two k-chunks
                const unsigned b_fields = b_lbo16 << 16;
                const unsigned kstep16 = 2u * b_lbo16;                                 // one K step of B in 16-byte units
                const unsigned a_lo_layer = A_DESC_FIELDS | ((smem_u32(S.act[li.in_buf]) + GUARD * 16) >> 4);
                SPX_TRACE_IF(leader, l, 0);
                bool skewed = false;
                if constexpr (NCTA == 2) {
                    if (l > 0) {   // 128 input channels: per-tile start (the stem's input is written by all warps at once)
                        issue_first_tap_skewed(S, li.taps, a_lo_layer, b_fields, kstep16, idesc, tmem_base, leader, stage, sphase, ephase);
                        skewed = true;
                    }
                }
                if (!skewed) {
                    for (int t = 0; t < MT; ++t) mbar_wait(&S.epi_done[t], ephase);   // (both CTAs:) inputs written, accumulators drained
                    tc_fence_after();
                    if constexpr (NCTA == 2) {   // bias slice = the first ring stage of the layer
#ifndef SPX_DBG_NO_TMA
                        mbar_wait(&S.full[stage], sphase);
                        mbar_wait(&S.peer_full[stage], sphase);
#endif
                        const unsigned bias_lo = b_fields | (smem_u32(S.wstage[stage]) >> 4);
                        const unsigned ones_lo = ((128u >> 4) << 16) | (smem_u32(S.ones) >> 4);
                        if (leader) {
#pragma unroll
                            for (int t = 0; t < MT; ++t) tc_mma_bias2(tmem_base + (unsigned)(t * 128), ones_lo, bias_lo, idesc);
#ifndef SPX_DBG_NO_TMA
                            tc_commit_t<NCTA>(&S.empty[stage]);
#endif
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; sphase ^= 1u; }
                    }
                }
                ephase ^= 1u;
                SPX_TRACE_IF(leader, l, 1);
                if (l == 0) issue_layer<NCTA, 1, 1>(S, li.taps, a_lo_layer, b_fields, kstep16, idesc, tmem_base, leader, stage, sphase);
                else issue_layer<NCTA, 2, CH / 32>(S, li.taps, a_lo_layer, b_fields, kstep16, idesc, tmem_base, leader, stage, sphase, skewed ? 1 : 0);
                if (skewed && li.taps == 1) {   // the 1x1 head conv ran entirely inside issue_first_tap_skewed
                    if (leader)
                        for (int t = 0; t < MT; ++t) tc_commit_t<NCTA>(&S.acc_full[t]);
                }
                SPX_TRACE_IF(leader, l, 2);
                __syncwarp();
            }
            if constexpr (NCTA == 2) {
                if (fused) {
                    // value layer Linear(1344 -> 256) (modules.py:104): D[hidden (128 per CTA) x 16 boards (8 per CTA)] += W1 * x^T,
                    // M=256, N=16, 84 K steps; A = weight stage (ring in buffer 0), B = the head activations (buffer 1).
                    // Four accumulators (K step mod 4, columns 0/16/32/48) keep consecutive MMAs independent.
                    SPX_TRACE_IF(leader, n_layers, 0);
                    for (int t = 0; t < MT; ++t) mbar_wait(&S.epi_done[t], ephase);   // head activations of both CTAs in place, accumulators drained
                    ephase ^= 1u;
                    tc_fence_after();
                    SPX_TRACE_IF(leader, n_layers, 1);
                    const unsigned idesc_fc = make_idesc(256, 16, F16);
                    const unsigned xv_lo = ((128u >> 4) << 16) | (smem_u32(S.act[1] + XV_OFF) >> 4);
                    for (int it = 0; it < fc_iters; ++it) {
                        mbar_wait(&S.fc_full[fstage], fphase);
                        mbar_wait(&S.fc_peer_full[fstage], fphase);
                        const unsigned a_lo = ((unsigned)(FC_KSTEP_BYTES / 2 >> 4) << 16) | (smem_u32(S.act[0] + fstage * FC_STAGE_BYTES) >> 4);
                        if (leader) {
#pragma unroll
                            for (int j = 0; j < 2; ++j) {
                                const int ks = 2 * it + j;
                                tc_mma_lo_acc2(tmem_base + (unsigned)(16 * (ks & 3)), a_lo + (unsigned)(j * (FC_KSTEP_BYTES >> 4)),
                                               xv_lo + (unsigned)(ks * (256 >> 4)), idesc_fc, ks >= 4 ? 1u : 0u);
                            }
                            tc_commit_t<NCTA>(&S.fc_empty[fstage]);
                        }
                        __syncwarp();
                        if (++fstage == FC_STAGES) { fstage = 0; fphase ^= 1u; }
                    }
                    if (leader) tc_commit_t<NCTA>(&S.fc_done);
                    SPX_TRACE_IF(leader, n_layers, 2);
                    __syncwarp();
                }
            }
        } else if (warp >= EPI_WARP0) {
            // ===================== epilogue warps (also write the stem input)
            const int et = tid - EPI_WARP0 * 32;          // 0..EPI_THREADS-1
            const int quarter = warp & 3, part = (warp - EPI_WARP0) >> 2;   // TMEM lane quarter, column part
            // "this CTA's activations for the next layer are in place": the leader's threads arrive locally, the peer CTA
            // meets on a named barrier and sends ONE remote arrive to the leader's barrier
            // "tile t of this CTA is ready for the next layer" (activations written and fenced, accumulator drained): one arrive
            // per warp on the LEADER CTA's barrier (remote from the peer CTA)
            auto signal_tile_done = [&](int t) {
                __syncwarp();
                if (lane == 0) {
                    if (crank == 0) mbar_arrive(&S.epi_done[t]);
                    else mbar_arrive_remote(&S.epi_done[t], 0);
                }
            };
            auto signal_epi_done = [&]() {   // all three tiles at once (after a CTA-wide barrier: every warp wrote to every tile)
                asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS));
                for (int t = 0; t < MT; ++t) signal_tile_done(t);
            };
            SPX_TT(0);
            const int ew = warp - EPI_WARP0;                       // epilogue warp index; warps ew < NB own game slot ew in fast mode
            const long long gbw = grp * NB + ew;
            const bool live_w = ENGINE && ew < NB && gbw < n_boards;
            if constexpr (ENGINE) {
                if (!shadow_all) {
                    // ---- fast mode: one simulation attempt of game ew between two passes (see the ENGINE comment above)
                    if (ew < NB) {
                        int emitted = 0;
                        if (lane == 0) { S.leaf2[ew][0] = 0ULL; S.leaf2[ew][1] = 0ULL; }
                        __syncwarp();
                        if (live_w) {
                            int st = ld_flag_uniform(&S.slot_status[ew]);
                            if (tick == 0) {   // first pass of the launch: the previous launch's outputs come from global memory
                                my_p = lane < n_act ? __ldcg(policy_out + gbw * n_act + lane) : 0.f;
                                v_next = __ldcg(value_out + gbw);
                            }
                            if (st == FS_DONE) { __threadfence(); st = FS_FAST; }   // the shadow warp is through with this game (leaf waiting, idle or parked)
                            if (E.cfg.reserved0 & 1) {   // timing experiment (SPX_DBG_FLAGS=1): no search at all, every pass re-evaluates the same leaf
                                if (lane == 0) { S.leaf2[ew][0] = __ldcg(E.leaf_own + gbw); S.leaf2[ew][1] = __ldcg(E.leaf_opp + gbw); }
                                emitted = 1;
                            } else
                            if (st == FS_FAST) {
                                const int fast_budget = (E.cfg.reserved0 >> 8) & 0xFF ? (E.cfg.reserved0 >> 8) & 0xFF : E.cfg.max_sims_per_tick;   // measured: 1 loses 5 % of the leaves, 8 costs 3 us more than 1
                                const int flags = engine_step<GAME, SPX_FAST_SOFT>(E, (int)gbw, my_p, v_next, fast_budget, 0, &S.leaf2[ew][0]);
                                emitted = flags & spx::ADV_EMITTED;
                                if (!flags) { __threadfence(); st = FS_TODO; }   // no leaf yet and not idle: the shadow warp goes on with it
                                if (lane == 0) st_volatile_s32(&S.slot_status[ew], st);
                            }
                        }
                        __syncwarp();
                        if (lane == 0) { S.own[ew] = S.leaf2[ew][0]; S.opp[ew] = S.leaf2[ew][1]; S.need[ew] = emitted ? 1 : 0; }
                    }
                } else if (et < NB) {
                    // ---- shadow-all mode: the shadow warp has produced this pass's leaves (job == pass)
                    while (ld_volatile_s32(&S.searches_done[0]) < pass + 1 || ld_volatile_s32(&S.searches_done[1]) < pass + 1) __nanosleep(100);
                    __threadfence();
                    const long long gb = grp * NB + et;
                    const bool ok = gb < n_boards;
                    S.own[et] = ok ? __ldcg(E.leaf_own + gb) : 0ULL;
                    S.opp[et] = ok ? __ldcg(E.leaf_opp + gb) : 0ULL;
                    S.need[et] = ok ? __ldcg(E.needs_eval + gb) : (unsigned char)0;
                }
            }
            SPX_TT(1);
            // work-conserving launch: one thread of the leader CTA (an epilogue warp that owns no game) draws the ticket for the
            // cluster's NEXT tick now and publishes the answer after the preprocess below -- the atomic's latency is off every
            // critical path, and every role warp of both CTAs finds the answer long before it gets to the top of that tick
            long long tk_prev = 0;
            bool tk_asked = false;
            if constexpr (ENGINE) {
                if (pass_budget >= 0 && ui == 0 && crank == 0 && ew == NB && lane == 0 && !go_final) {
                    tk_prev = (long long)atomicAdd(E.ticks + 1, (unsigned long long)U);
                    tk_asked = true;
                }
            }
            if (fused && !first_unit) {
                // the previous unit's FC ring / head activations overwrote the buffers: the zero guard rows (never written by
                // an epilogue) must read as zero again; every other row is rewritten before it is read
                const int buf = et >> 8, chunk = (et >> 4) & 15, gr = et & 15;   // 2 buffers x 16 chunks x 16 guard rows = 512 threads
                const int row = gr < GUARD ? gr : ROWS + gr;
                *reinterpret_cast<uint4*>(S.act[buf] + chunk * CHUNK_BYTES + row * 16) = make_uint4(0, 0, 0, 0);
            }
            if constexpr (!ENGINE) {
                if (et < NB) {
                    const long long gb = grp * NB + et;
                    S.own[et] = gb < n_boards ? own_g[gb] : 0ULL;
                    S.opp[et] = gb < n_boards ? opp_g[gb] : 0ULL;
                }
            }
            if constexpr (NCTA == 1) {
                if (et < CH) { S.bias[0][et] = __ldg(bias_all + et); S.bias[1][et] = __ldg(bias_all + CH + et); }
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS));
            SPX_TT(2);
            if constexpr (NCTA == 1) {
                store_bias_to_tmem(tmem_base + ((unsigned)(quarter * 32) << 16) + (unsigned)(part * (CH / EPI_SPLIT)), S.bias[0] + part * (CH / EPI_SPLIT));
                tc_fence_before();
            }
            // preprocess (modules.py:115-125): planes (empty, own, enemy) -> channels 0..2 of buffer 0, channels 3..15 zero
            for (int row = et; row < ROWS; row += EPI_THREADS) {
                int board, cell, bit;
                const bool real = row_is_cell(game, row, board, cell, bit);
                unsigned o = 0, e = 0;
                if (real) { o = (unsigned)((S.own[board] >> bit) & 1ULL); e = (unsigned)((S.opp[board] >> bit) & 1ULL); }
                const unsigned emp = real ? (1u - o - e) : 0u;
                const unsigned one = F16 ? 0x3C00u : 0x3F80u;  // 1.0 in the activation type
                uint4 v0 = make_uint4((emp ? one : 0u) | ((o ? one : 0u) << 16), e ? one : 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(S.act[0] + (GUARD + row) * 16) = v0;
                *reinterpret_cast<uint4*>(S.act[0] + CHUNK_BYTES + (GUARD + row) * 16) = make_uint4(0, 0, 0, 0);
            }
            fence_proxy_async();
            signal_epi_done();
            SPX_TT(3);
            if constexpr (ENGINE) {
                if (tk_asked) {
                    const bool ok = tk_prev + U <= pass_budget;
                    const int w = ok ? tick + 2 : ((tick + 1) | GO_FINAL);
                    go_final = !ok;
                    st_volatile_s32(&S.go, w);
                    if constexpr (NCTA == 2) st_shared_remote_s32(&S.go, 1u, w);
                }
            }
            // per-tile row bookkeeping is layer independent
            bool real_t[MT]; int board_t[MT], cell_t[MT];
#pragma unroll
            for (int t = 0; t < MT; ++t) { int bit; real_t[t] = row_is_cell(game, t * 128 + quarter * 32 + lane, board_t[t], cell_t[t], bit); }
            for (int l = 0; l < n_layers; ++l) {
                const LayerInfo li = layer_info(l, n_layers);
                // stage the bias of layer l+2 while this layer's MMAs are still running (layer l+1's is already visible)
                if constexpr (NCTA == 1) {
                    if (l + 2 < n_layers && et < CH) S.bias[(l + 2) % 3][et] = __ldg(bias_all + (size_t)(l + 2) * CH + et);
                    asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS));   // published one layer before its use
                }
                SPX_TRACE_IF(et == 0, l, 3);
                mbar_wait_backoff(&S.acc_full[0], lphase);
                if (li.out_buf < 0) { mbar_wait(&S.acc_full[1], lphase); mbar_wait(&S.acc_full[2], lphase); }   // head layer: all tiles
                tc_fence_after();
                SPX_TRACE_IF(et == 0, l, 4);
                if (l == 0) SPX_TT(4);
                if (li.out_buf >= 0) {
                    // trunk layer: this warp owns 128/EPI_SPLIT = 32 columns of its 32 rows, for each of the 3 row tiles
                    const int ch0 = part * (CH / EPI_SPLIT);
                    const unsigned tcol = tmem_base + ((unsigned)(quarter * 32) << 16) + (unsigned)ch0;
                    unsigned v[2][32];
                    tc_ld32_nowait(tcol, v[0]);
#pragma unroll
                    for (int t = 0; t < MT; ++t) {
                        tc_wait_ld();
                        SPX_TRACE_IF(et == 0, l, 5 + t);
                        if (t + 1 < MT) {
                            mbar_wait(&S.acc_full[t + 1], lphase);   // complete long ago (a few MMAs after tile t)
                            tc_fence_after();
                            tc_ld32_nowait(tcol + (unsigned)((t + 1) * 128), v[(t + 1) & 1]);
                        }
                        const int row = t * 128 + quarter * 32 + lane;
                        unsigned char* obase = S.act[li.out_buf] + (GUARD + row) * 16 + (ch0 >> 3) * CHUNK_BYTES;
#ifndef SPX_DBG_SKIP_EPI
                        if (!real_t[t]) {   // padding / guard cell: must read as zero in the next layer
#pragma unroll
                            for (int g8 = 0; g8 < 4; ++g8) *reinterpret_cast<uint4*>(obase + g8 * CHUNK_BYTES) = make_uint4(0, 0, 0, 0);
                        } else {
                            const unsigned* vv = v[t & 1];
#pragma unroll
                            for (int g8 = 0; g8 < 4; ++g8) {
                                float y[8];   // accumulator already contains the folded-BN bias
#pragma unroll
                                for (int k = 0; k < 8; ++k) y[k] = __uint_as_float(vv[g8 * 8 + k]);
                                uint4* dst = reinterpret_cast<uint4*>(obase + g8 * CHUNK_BYTES);
                                if (li.residual) {  // out += identity (modules.py:37), identity lives in the output buffer
                                    const uint4 idv = *dst;
                                    const unsigned iw[4] = {idv.x, idv.y, idv.z, idv.w};
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        const float2 id2 = unpack2<F16>(iw[k]);
                                        y[2 * k] += id2.x;
                                        y[2 * k + 1] += id2.y;
                                    }
                                }
                                unsigned pk[4];
#pragma unroll
                                for (int k = 0; k < 4; ++k) pk[k] = relu_pack2<F16>(y[2 * k], y[2 * k + 1]);
                                *dst = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                            }
                        }
#endif
                        if constexpr (NCTA == 2) {   // tile t is complete: the next layer may start on it (issue_first_tap_skewed)
                            tc_fence_before();
                            fence_proxy_async();
                            signal_tile_done(t);
                        }
                    }
                    SPX_TRACE_IF(et == 0, l, 8);
                    if constexpr (NCTA == 1) {
                        if (l + 1 < n_layers) store_bias_to_tmem(tcol, S.bias[(l + 1) % 3] + ch0);
                    }
                    SPX_TRACE_IF(et == 0, l, 9);
                } else {
                    // fused policy/value 1x1 head conv + BN + ReLU (modules.py:97,102); 64 columns: each warp owns 64/EPI_SPLIT = 16.
                    // Separate heads kernel: fp32 [board][ch*42 + cell] to global memory.  Fused FC heads: policy channels as fp32
                    // [board][k] and value channels as bf16 in the K-major core-matrix layout [k-chunk][8 boards][8] (board 7 =
                    // zero padding) in the dead activation buffer 1, k = channel*42 + cell (the flatten order of modules.py:98,103).
                    const int ch0 = part * (HEAD_CH / EPI_SPLIT);
                    if (fused && et < flat / 8) *reinterpret_cast<uint4*>(S.act[1] + XV_OFF + et * 128 + 7 * 16) = make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int t = 0; t < MT; ++t) {
                        unsigned v[16];
                        tc_ld16(tmem_base + ((unsigned)(quarter * 32) << 16) + (unsigned)(t * 128 + ch0), v);
                        const long long gb = grp * NB + board_t[t];
                        if (!fused) {
                            if (real_t[t] && gb < n_boards) {
                                float* ob = head_out + (size_t)gb * (HEAD_CH * cells) + cell_t[t];
#pragma unroll
                                for (int k = 0; k < 16; ++k) ob[(size_t)(ch0 + k) * cells] = fmaxf(__uint_as_float(v[k]), 0.f);
                            }
                        } else if (real_t[t]) {
                            if (ch0 < 32) {
                                float* xp = reinterpret_cast<float*>(S.act[1] + XP_OFF) + board_t[t] * flat + cell_t[t];
#pragma unroll
                                for (int k = 0; k < 16; ++k) xp[(ch0 + k) * cells] = fmaxf(__uint_as_float(v[k]), 0.f);
                            } else {
                                unsigned char* xv = S.act[1] + XV_OFF + board_t[t] * 16;
#pragma unroll
                                for (int k = 0; k < 16; ++k) {
                                    const int kk = (ch0 - 32 + k) * cells + cell_t[t];
                                    const float xr = fmaxf(__uint_as_float(v[k]), 0.f);
                                    if constexpr (F16) *reinterpret_cast<__half*>(xv + (kk >> 3) * 128 + (kk & 7) * 2) = __float2half_rn(fminf(xr, 65504.f));
                                    else *reinterpret_cast<__nv_bfloat16*>(xv + (kk >> 3) * 128 + (kk & 7) * 2) = __float2bfloat16_rn(xr);
                                }
                            }
                        }
                        __syncwarp();
                    }
                }
                tc_fence_before();
                fence_proxy_async();
                lphase ^= 1u;
                SPX_TRACE_IF(et == 0, l, 10);
                if (NCTA == 2 && li.out_buf >= 0) { }                        // trunk layer of the SM-pair kernel: signalled tile by tile
                else if (l + 1 < n_layers || fused) signal_epi_done();
                else asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS));
                SPX_TRACE_IF(et == 0, l, 11);
            }
            if constexpr (NCTA == 2) {
                if (fused) {
                    SPX_TT(5);
                    float* scr = reinterpret_cast<float*>(S.act[1] + FCS_OFF);   // logits [7][9] | partial [4][16] at 64 | own [16] at 128 | peer [16] at 144
                    // ---- policy head Linear(1344 -> A) + softmax (modules.py:99-100) in fp32 on the CUDA cores, while the tensor
                    // pipe runs the value layer: one warp per (board, action) dot product
                    const float* xp = reinterpret_cast<const float*>(S.act[1] + XP_OFF);
                    SPX_TRACE_IF(et == 0, n_layers, 3);
                    for (int task = warp - EPI_WARP0; task < NB * n_act; task += EPI_WARPS) {
                        const int b = task / n_act, a = task - b * n_act;
                        float acc = 0.f;
                        for (int k = lane; k < flat; k += 32) acc = fmaf(xp[b * flat + k], __ldg(polw + a * flat + k), acc);
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
                        if (lane == 0) scr[b * 9 + a] = acc + __ldg(polb + a);
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS));
                    SPX_TT(6);
                    bool had_w = false;              // ENGINE: game slot `ew` of this CTA was evaluated by this pass
                    if constexpr (ENGINE) {
                        had_w = live_w && S.need[ew] != 0;
                        if (et == 0) mbar_expect_tx(&S.xbar, 32u);   // this pass's exchange: 8 floats from the peer CTA
                        if (had_w) {   // softmax of board ew by its own warp: the same sequential fp32 arithmetic in every lane
                            float m = scr[ew * 9];
                            for (int a = 1; a < n_act; ++a) m = fmaxf(m, scr[ew * 9 + a]);
                            float z = 0.f;
                            for (int a = 0; a < n_act; ++a) z += expf(scr[ew * 9 + a] - m);
                            my_p = lane < n_act ? expf(scr[ew * 9 + lane] - m) / z : 0.f;
                            if (lane < n_act) policy_out[gbw * n_act + lane] = my_p;
                        }
                    } else if (et < NB) {
                        const long long gb = grp * NB + et;
                        if (gb < n_boards && ld_need(gb)) {
                            float m = scr[et * 9];
                            for (int a = 1; a < n_act; ++a) m = fmaxf(m, scr[et * 9 + a]);
                            float e[SPX_MAX_ACTIONS], z = 0.f;
                            for (int a = 0; a < n_act; ++a) { e[a] = expf(scr[et * 9 + a] - m); z += e[a]; }
                            for (int a = 0; a < n_act; ++a) policy_out[gb * n_act + a] = e[a] / z;
                        }
                    }
                    // ---- value head: relu(W1 x + b1) . w2 (+ b2, tanh after the two halves of the hidden layer met)
                    SPX_TRACE_IF(et == 0, n_layers, 4);
                    mbar_wait_backoff(&S.fc_done, dphase);
                    dphase ^= 1u;
                    tc_fence_after();
                    SPX_TRACE_IF(et == 0, n_layers, 5);
                    SPX_TT(7);
                    if (part == 0) {
                        const unsigned tl = tmem_base + ((unsigned)(quarter * 32) << 16);
                        unsigned v0[16], v1[16], v2[16], v3[16];
                        tc_ld16(tl, v0); tc_ld16(tl + 16, v1); tc_ld16(tl + 32, v2); tc_ld16(tl + 48, v3);
                        const int j = (int)crank * 128 + quarter * 32 + lane;      // hidden unit of this TMEM lane
                        const float b1 = __ldg(fc_b1 + j), w2 = __ldg(fc_w2 + j);
#pragma unroll
                        for (int c = 0; c < 16; ++c) {
                            const float h = (__uint_as_float(v0[c]) + __uint_as_float(v1[c])) + (__uint_as_float(v2[c]) + __uint_as_float(v3[c])) + b1;
                            float y = fmaxf(h, 0.f) * w2;
#pragma unroll
                            for (int off = 16; off > 0; off >>= 1) y += __shfl_xor_sync(0xffffffffu, y, off);
                            if (lane == 0) scr[64 + quarter * 16 + c] = y;
                        }
                    }
                    tc_fence_before();
                    asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS));
                    SPX_TT(8);
                    if constexpr (ENGINE) {
                        // column c: boards 0..7 of CTA 0, 8..15 of CTA 1; every CTA needs both halves of the hidden layer for ITS boards:
                        // the half sums of the peer's boards go to the peer (st.async completes the peer's xbar), mine are summed here
                        if (et < 8) {
                            const int c = (int)(crank ^ 1u) * 8 + et;
                            const float tot = (scr[64 + c] + scr[64 + 16 + c]) + (scr[64 + 32 + c] + scr[64 + 48 + c]);
                            st_async_remote_f32(&S.xval[et], &S.xbar, crank ^ 1u, tot);
                        }
                        if (had_w) {
                            const int c = (int)crank * 8 + ew;
                            const float mine = (scr[64 + c] + scr[64 + 16 + c]) + (scr[64 + 32 + c] + scr[64 + 48 + c]);
                            mbar_wait(&S.xbar, (unsigned)pass & 1u);
                            const float peer = S.xval[ew];
                            const float lo = crank == 0 ? mine : peer, hi = crank == 0 ? peer : mine;     // hidden 0..127, 128..255
                            v_next = tanhf((lo + hi) + __ldg(fc_b2));                                     // Linear(256 -> 1) + tanh (modules.py:105)
                            if (lane == 0) value_out[gbw] = v_next;
                        }
                    } else
                    if (et < 16) {   // column c: boards 0..7 of CTA 0, 8..15 of CTA 1; both CTAs need both halves of the hidden layer
                        const float tot = (scr[64 + et] + scr[64 + 16 + et]) + (scr[64 + 32 + et] + scr[64 + 48 + et]);
                        scr[128 + et] = tot;
                        st_shared_remote_f32(&scr[144 + et], crank ^ 1u, tot);
                    }
                    SPX_TRACE_IF(et == 0, n_layers, 6);
                    SPX_TT(9);
                }
            } else {
                (void)dphase;
            }
            asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS));
            SPX_TT(10);
            if constexpr (ENGINE) {
                // every output of this pass is written: the shadow warp may consume them (shadow-all mode)
                if (shadow_all && et == 0) { __threadfence(); st_volatile_s32(&S.passes_done, pass + 1); }
            }
        }
        // non-elected lanes of warps 0-2 and warp 3 fall through; ring/phase state persists in the role warps
        if constexpr (!ENGINE) {
        __syncthreads();   // unit boundary: accumulators drained, buffers reusable
        if constexpr (NCTA == 2) {
            cluster_sync_all();
            if (fused && warp == EPI_WARP0 && lane < NB) {   // both halves of the hidden layer are in: Linear(256 -> 1) + tanh (modules.py:105)
                const float* scr = reinterpret_cast<const float*>(S.act[1] + FCS_OFF);
                const long long gb = grp * NB + lane;
                const int c = (int)crank * 8 + lane;
                const float lo = crank == 0 ? scr[128 + c] : scr[144 + c], hi = crank == 0 ? scr[144 + c] : scr[128 + c];   // hidden 0..127, 128..255
                if (gb < n_boards && ld_need(gb)) value_out[gb] = tanhf((lo + hi) + __ldg(fc_b2));
            }
            if (fused) cluster_sync_all();   // the scratch may be overwritten by the peer in the next unit only after it was read
        }
        }
        first_unit = false;
    }
    }   // tick
    if constexpr (ENGINE) {
        if (tid == EPI_WARP0 * 32) {
            st_volatile_s32(&S.quit, 1);                                  // fast mode: the shadow warp stops after the game it is on
            if (blockIdx.x == 0) atomicAdd(E.ticks, (unsigned long long)n_ticks);
        }
    }

    if (tid == 0) SPX_TRACE(62, 2);
#ifdef SPX_DBG_TRACE
    if (tid == 0 && blockIdx.x < 160) g_trace[1024 + 3 * blockIdx.x + 1] = (long long)globaltimer_ns();
#endif
    tc_fence_before();
    __syncthreads();
    if constexpr (NCTA == 2) {
        cluster_sync_all();
        if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    } else {
        if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ fully connected heads
// policy: Linear(1344 -> A) + softmax (modules.py:99-100); value: Linear(1344 -> 256) + ReLU + Linear(256 -> 1) + tanh
// (modules.py:104-105).  grid = (boards/16, 2): blockIdx.y == 0 computes the value head, == 1 the policy head.
// The 1344x256 value layer (352 MMAC for 1024 boards, 0.07 % of the network) runs on warp-level mma.sync m16n8k16
// (bf16 in, fp32 accumulate): warp w owns hidden units [32w, 32w+32); its weights are streamed through shared memory
// in 64-wide K chunks with a double-buffered cp.async pipeline.  The policy layer stays fp32 on CUDA cores.
constexpr int HB = 16;
constexpr int XS_STRIDE = FLAT + 8;        // bf16 elements per staged activation row (+8: conflict-free fragment loads)
constexpr int KC = 64;                     // K chunk of the weight pipeline
constexpr int WS_STRIDE = KC + 8;          // bf16 elements per staged weight row
constexpr int HEADS_SMEM = HB * XS_STRIDE * 2 + 2 * FC_HIDDEN * WS_STRIDE * 2 + HB * FC_HIDDEN * 4;

template <bool F16>
__device__ __forceinline__ void mma_16816(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
    if constexpr (F16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}

template <bool F16>
__global__ void __launch_bounds__(256) heads_kernel(const float* __restrict__ head_in, const unsigned char* __restrict__ needs,
                                                    long long n_boards, int A, const float* __restrict__ pol_w,
                                                    const float* __restrict__ pol_b, const __nv_bfloat16* __restrict__ w1,
                                                    const float* __restrict__ b1, const float* __restrict__ w2, const float* __restrict__ b2,
                                                    float* __restrict__ policy, float* __restrict__ value) {
    extern __shared__ __align__(16) unsigned char hsm[];
    const long long b0 = (long long)blockIdx.x * HB;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    bool any = needs == nullptr;
    if (!any) for (int b = 0; b < HB; ++b) if (b0 + b < n_boards && needs[b0 + b]) any = true;
    if (!any) return;

    if (blockIdx.y == 1) {
        // ---------------- policy head: thread = (board, 1/16 of K); fp32
        const int bl = tid >> 4, part = tid & 15;
        const long long gb = b0 + bl;
        float d[SPX_MAX_ACTIONS];
#pragma unroll
        for (int a = 0; a < SPX_MAX_ACTIONS; ++a) d[a] = 0.f;
        if (gb < n_boards) {
            const float* pin = head_in + (size_t)gb * (HEAD_CH * CELLS);
#pragma unroll 4
            for (int k = part; k < FLAT; k += 16) {
                const float x = pin[k];
#pragma unroll
                for (int a = 0; a < SPX_MAX_ACTIONS; ++a) if (a < A) d[a] = fmaf(x, __ldg(pol_w + (size_t)a * FLAT + k), d[a]);
            }
        }
        float m = -INFINITY;
#pragma unroll
        for (int a = 0; a < SPX_MAX_ACTIONS; ++a) {
            for (int off = 8; off > 0; off >>= 1) d[a] += __shfl_xor_sync(0xffffffffu, d[a], off);
            d[a] = a < A ? d[a] + pol_b[a] : -INFINITY;
            m = fmaxf(m, d[a]);
        }
        float z = 0.f;
#pragma unroll
        for (int a = 0; a < SPX_MAX_ACTIONS; ++a) { d[a] = a < A ? expf(d[a] - m) : 0.f; z += d[a]; }
        if (part == 0 && gb < n_boards) {
#pragma unroll
            for (int a = 0; a < SPX_MAX_ACTIONS; ++a) if (a < A) policy[gb * A + a] = d[a] / z;
        }
        return;
    }

    // ---------------- value head
    __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(hsm);                                   // [HB][XS_STRIDE]
    __nv_bfloat16* ws = reinterpret_cast<__nv_bfloat16*>(hsm + HB * XS_STRIDE * 2);              // [2][256][WS_STRIDE]
    float* hid = reinterpret_cast<float*>(hsm + HB * XS_STRIDE * 2 + 2 * FC_HIDDEN * WS_STRIDE * 2);  // [HB][256]
    auto load_chunk = [&](int c, int buf) {   // w1[n][c*64 .. +64) -> ws[buf][n][0..64): 8 x 16 B per row, 2048 copies / 256 threads
        __nv_bfloat16* dst = ws + (size_t)buf * FC_HIDDEN * WS_STRIDE;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int idx = tid + i * 256, n = idx >> 3, seg = idx & 7;
            cp_async16(dst + n * WS_STRIDE + seg * 8, w1 + (size_t)n * FLAT + c * KC + seg * 8);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    load_chunk(0, 0);
    for (int i = tid; i < HB * (FLAT / 4); i += 256) {
        const int b = i / (FLAT / 4), k4 = (i - b * (FLAT / 4)) * 4;
        float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b0 + b < n_boards) x = *reinterpret_cast<const float4*>(head_in + (size_t)(b0 + b) * (HEAD_CH * CELLS) + FLAT + k4);
        unsigned plo, phi;   // the 16-bit type of the tower's weight stream (w1 is stored in it)
        if constexpr (F16) {
            __half2 lo = __floats2half2_rn(fminf(x.x, 65504.f), fminf(x.y, 65504.f)), hi = __floats2half2_rn(fminf(x.z, 65504.f), fminf(x.w, 65504.f));
            plo = *reinterpret_cast<unsigned*>(&lo); phi = *reinterpret_cast<unsigned*>(&hi);
        } else {
            __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y), hi = __floats2bfloat162_rn(x.z, x.w);
            plo = *reinterpret_cast<unsigned*>(&lo); phi = *reinterpret_cast<unsigned*>(&hi);
        }
        *reinterpret_cast<uint2*>(xs + b * XS_STRIDE + k4) = make_uint2(plo, phi);
    }
    const int r = lane >> 2, kq = (lane & 3) * 2, n0 = warp * 32;
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
    constexpr int NCHUNK = FLAT / KC;   // 21
    for (int c = 0; c < NCHUNK; ++c) {
        if (c + 1 < NCHUNK) { load_chunk(c + 1, (c + 1) & 1); asm volatile("cp.async.wait_group 1;" ::: "memory"); }
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        const __nv_bfloat16* wb = ws + (size_t)(c & 1) * FC_HIDDEN * WS_STRIDE + (n0 + r) * WS_STRIDE + kq;
        const __nv_bfloat16* xr0 = xs + r * XS_STRIDE + c * KC + kq;
        const __nv_bfloat16* xr1 = xr0 + 8 * XS_STRIDE;
#pragma unroll
        for (int kk = 0; kk < KC; kk += 16) {
            const unsigned a0 = *reinterpret_cast<const unsigned*>(xr0 + kk), a1 = *reinterpret_cast<const unsigned*>(xr1 + kk);
            const unsigned a2 = *reinterpret_cast<const unsigned*>(xr0 + kk + 8), a3 = *reinterpret_cast<const unsigned*>(xr1 + kk + 8);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const __nv_bfloat16* wp = wb + nt * 8 * WS_STRIDE + kk;
                mma_16816<F16>(acc[nt], a0, a1, a2, a3, *reinterpret_cast<const unsigned*>(wp), *reinterpret_cast<const unsigned*>(wp + 8));
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int n = n0 + nt * 8 + kq;
        const float bia0 = b1[n], bia1 = b1[n + 1];
        hid[r * FC_HIDDEN + n] = fmaxf(acc[nt][0] + bia0, 0.f);
        hid[r * FC_HIDDEN + n + 1] = fmaxf(acc[nt][1] + bia1, 0.f);
        hid[(r + 8) * FC_HIDDEN + n] = fmaxf(acc[nt][2] + bia0, 0.f);
        hid[(r + 8) * FC_HIDDEN + n + 1] = fmaxf(acc[nt][3] + bia1, 0.f);
    }
    __syncthreads();
    for (int bl = warp; bl < HB; bl += 8) {   // value = tanh(hid . w2 + b2)
        const long long gb = b0 + bl;
        if (gb >= n_boards) continue;
        float sacc = 0.f;
        for (int j = lane; j < FC_HIDDEN; j += 32) sacc = fmaf(hid[bl * FC_HIDDEN + j], w2[j], sacc);
        for (int off = 16; off > 0; off >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, off);
        if (lane == 0) value[gb] = tanhf(sacc + b2[0]);
    }
}

}  // namespace tower
}  // namespace spx

// ================================================================================================== C ABI
struct spx_tower {
    int game, num_blocks, n_layers, A, ncta, fused, f16;
    size_t off_bias, off_polw, off_polb, off_w1t, off_b1, off_w2, off_b2, blob_bytes;
    unsigned char* blob;    // device copy of the packed weights
    float* head_buf;        // [capacity][64*42] fp32 head-conv activations
    long long capacity;
    int sm_count;
    unsigned version;       // unique per spx_tower_load of any tower: the tag of the engines' evaluation-cache entries
};

using namespace spx::tower;

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

extern "C" {

struct TowerLayout { int A, flat; size_t conv, off_bias, off_polw, off_polb, off_w1t, off_b1, off_w2, off_b2, bytes; };
/* blob layout of nets.pack_tower_blob: the kernel's weight stream (per layer bias slice + conv weights, then the value layer
 * as ring stages), then fp32 biases, policy Linear, bf16 value Linear 1, fp32 b1, w2, b2 -- every piece 256-byte aligned */
static TowerLayout tower_layout(int game, int num_blocks) {
    TowerLayout L;
    const int cells = game == SPX_GAME_TICTACTOE ? 9 : CELLS, n_layers = 2 * num_blocks + 2;
    L.A = game == SPX_GAME_TICTACTOE ? 9 : 7;
    L.flat = 32 * cells;
    L.conv = (size_t)9 * KSTEP_BYTES + (size_t)num_blocks * 2 * 72 * KSTEP_BYTES + (size_t)8 * (2 * HEAD_CH * 16) + bias_slices_bytes(num_blocks) +
             (size_t)(L.flat / 16) * 2 * FC_KSTEP_BYTES;
    size_t off = align_up(L.conv, 256);
    L.off_bias = off; off = align_up(off + (size_t)n_layers * CH * 4, 256);
    L.off_polw = off; off = align_up(off + (size_t)L.A * L.flat * 4, 256);
    L.off_polb = off; off = align_up(off + 16 * 4, 256);
    L.off_w1t = off; off = align_up(off + (size_t)L.flat * FC_HIDDEN * 2, 256);
    L.off_b1 = off; off = align_up(off + FC_HIDDEN * 4, 256);
    L.off_w2 = off; off = align_up(off + FC_HIDDEN * 4, 256);
    L.off_b2 = off; off = align_up(off + 16, 256);
    L.bytes = off;
    return L;
}

int64_t spx_tower_blob_bytes(int32_t game, int32_t num_blocks) {
    if ((game != SPX_GAME_CONNECT4 && game != SPX_GAME_TICTACTOE) || num_blocks < 0) return -1;
    return (int64_t)tower_layout(game, num_blocks).bytes;
}

int spx_tower_create(int32_t game, int32_t num_blocks, spx_tower** out) {
    if (!out) return spx::set_err(SPX_E_ARG, "spx_tower_create: null out%s", "");
    if (game != SPX_GAME_CONNECT4 && game != SPX_GAME_TICTACTOE) return spx::set_err(SPX_E_ARG, "spx_tower_create: unknown game%s", "");
    if (num_blocks < 1 || num_blocks > 64) return spx::set_err(SPX_E_ARG, "spx_tower_create: num_blocks out of range%s", "");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return spx::set_err(SPX_E_CUDA, "spx_tower_create: no CUDA device (there is no CPU fallback)%s", "");
    spx_tower* t = new (std::nothrow) spx_tower();
    if (!t) return spx::set_err(SPX_E_ARG, "spx_tower_create: out of host memory%s", "");
    memset(t, 0, sizeof(*t));
    const TowerLayout L = tower_layout(game, num_blocks);
    t->game = game; t->num_blocks = num_blocks; t->n_layers = 2 * num_blocks + 2; t->A = L.A;
    {
        const char* e = getenv("SPX_TOWER_NCTA");   // 2 (default): SM-pair kernel (cta_group::2); 1: single-CTA kernel
        t->ncta = (e && e[0] == '1') ? 1 : 2;
        const char* f = getenv("SPX_TOWER_FUSED_HEADS");   // 1 (default, SM-pair kernel only): FC heads inside the tower kernel; 0: separate heads kernel
        t->fused = (t->ncta == 2 && !(f && f[0] == '0')) ? 1 : 0;
        // activation / weight type of the conv trunk and the fused value layer: fp16 (default: 11 significant bits, the type the
        // reference's own GPU path computes in -- torch.cuda.amp.autocast, inference_worker.py:117) or bf16 (SPX_TOWER_DTYPE=bf16).
        // Same tcgen05 kind::f16 instruction and rate either way; fp16 needs the SM-pair kernel.
        const char* d = getenv("SPX_TOWER_DTYPE");
        t->f16 = (t->ncta == 2 && !(d && (d[0] == 'b' || d[0] == 'B'))) ? 1 : 0;
    }
    if (game == SPX_GAME_TICTACTOE && !t->fused) {
        delete t;
        return spx::set_err(SPX_E_ARG, "spx_tower_create: the 3x3 ResidualTower runs on the SM-pair kernel with fused heads only (unset SPX_TOWER_NCTA / SPX_TOWER_FUSED_HEADS)%s", "");
    }
    t->off_bias = L.off_bias; t->off_polw = L.off_polw; t->off_polb = L.off_polb; t->off_w1t = L.off_w1t;
    t->off_b1 = L.off_b1; t->off_w2 = L.off_w2; t->off_b2 = L.off_b2; t->blob_bytes = L.bytes;
    SPX_CUDA_T(cudaMalloc((void**)&t->blob, t->blob_bytes));
    int dev = 0;
    SPX_CUDA_T(cudaGetDevice(&dev));
    SPX_CUDA_T(cudaDeviceGetAttribute(&t->sm_count, cudaDevAttrMultiProcessorCount, dev));
    SPX_CUDA_T(cudaFuncSetAttribute(tower_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemT<1>)));
    SPX_CUDA_T(cudaFuncSetAttribute(tower_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemT<2>)));
    SPX_CUDA_T(cudaFuncSetAttribute(tower_kernel<2, SPX_GAME_TICTACTOE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemT<2>)));
    SPX_CUDA_T(cudaFuncSetAttribute(tower_kernel<2, SPX_GAME_CONNECT4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemT<2>)));
    SPX_CUDA_T(cudaFuncSetAttribute(tower_kernel<2, SPX_GAME_TICTACTOE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemT<2>)));
    SPX_CUDA_T(cudaFuncSetAttribute(tower_kernel<2, SPX_GAME_CONNECT4, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemT<2>)));
    SPX_CUDA_T(cudaFuncSetAttribute(tower_kernel<2, SPX_GAME_TICTACTOE, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemT<2>)));
    SPX_CUDA_T(cudaFuncSetAttribute(tower_kernel<2, SPX_GAME_CONNECT4, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemT<2>)));
    SPX_CUDA_T(cudaFuncSetAttribute(tower_kernel<2, SPX_GAME_TICTACTOE, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SmemT<2>)));
    SPX_CUDA_T(cudaFuncSetAttribute(heads_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, HEADS_SMEM));
    SPX_CUDA_T(cudaFuncSetAttribute(heads_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, HEADS_SMEM));
    *out = t;
    return 0;
}

/* 1 or 2: which weight-slice layout spx_tower_load expects (2 = output channels split across the SM pair) */
int spx_tower_ncta(spx_tower* t) { return t ? t->ncta : 0; }
/* 1: the conv trunk and the fused value layer compute in fp16 (default), 0: bf16 -- the type spx_tower_load expects the weight
 * stream in (nets.pack_tower_blob(dtype=...)) */
int spx_tower_f16(spx_tower* t) { return t ? t->f16 : 0; }
uint32_t spx_tower_version(spx_tower* t) { return t ? t->version : 0u; }
/* 1 when the fully connected heads run inside the tower kernel (SM-pair kernel, default), 0 when heads_kernel follows it */
int spx_tower_fused_heads(spx_tower* t) { return t ? t->fused : 0; }

int spx_tower_destroy(spx_tower* t) {
    if (!t) return 0;
    if (t->blob) cudaFree(t->blob);
    if (t->head_buf) cudaFree(t->head_buf);
    delete t;
    return 0;
}

/* copies a packed weight blob (device pointer, layout of nets.pack_tower_blob / spx_tower_blob_bytes) */
int spx_tower_load(spx_tower* t, const void* dev_blob, int64_t bytes, void* stream) {
    if (!t || !dev_blob) return spx::set_err(SPX_E_ARG, "spx_tower_load: null argument%s", "");
    if ((size_t)bytes != t->blob_bytes) return spx::set_err(SPX_E_ARG, "spx_tower_load: blob size mismatch%s", "");
    SPX_CUDA_T(cudaMemcpyAsync(t->blob, dev_blob, t->blob_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    static std::atomic<unsigned> g_version{0};
    t->version = ++g_version;   // new weights: cached evaluations of older versions never match again
    return 0;
}

static int tower_forward_impl(spx_tower* t, const uint64_t* own, const uint64_t* opp, const uint8_t* needs_eval, int64_t n,
                              float* policy, float* value, cudaStream_t st, cudaEvent_t e0, cudaEvent_t e1, cudaEvent_t e2) {
    if (!t || !own || !opp || !policy || !value) return spx::set_err(SPX_E_ARG, "spx_tower_forward: null argument%s", "");
    if (n <= 0) return 0;
    if (!t->fused && n > t->capacity) {   // the fused-heads kernel keeps the head activations in shared memory
        if (t->head_buf) { SPX_CUDA_T(cudaStreamSynchronize(st)); SPX_CUDA_T(cudaFree(t->head_buf)); t->head_buf = nullptr; }
        SPX_CUDA_T(cudaMalloc((void**)&t->head_buf, (size_t)n * HEAD_CH * CELLS * sizeof(float)));
        t->capacity = n;
    }
    const long long groups = (n + NB - 1) / NB;
    if (e0) SPX_CUDA_T(cudaEventRecord(e0, st));
    if (t->ncta == 2) {
        const long long pairs = (groups + 1) / 2, max_pairs = t->sm_count / 2;
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3((unsigned)(2 * (pairs < max_pairs ? pairs : max_pairs)));
        cfg.blockDim = dim3(NUM_THREADS);
        cfg.dynamicSmemBytes = sizeof(SmemT<2>);
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        auto kern = t->f16 ? (t->game == SPX_GAME_TICTACTOE ? tower_kernel<2, SPX_GAME_TICTACTOE, false, true> : tower_kernel<2, SPX_GAME_CONNECT4, false, true>)
                           : (t->game == SPX_GAME_TICTACTOE ? tower_kernel<2, SPX_GAME_TICTACTOE> : tower_kernel<2, SPX_GAME_CONNECT4>);
        SPX_CUDA_T(cudaLaunchKernelEx(&cfg, kern, (const unsigned long long*)own, (const unsigned long long*)opp, needs_eval,
                                      (long long)n, t->n_layers, (const unsigned char*)t->blob, (const float*)(t->blob + t->off_bias), t->head_buf,
                                      t->fused, (const float*)(t->blob + t->off_polw), (const float*)(t->blob + t->off_polb), (const float*)(t->blob + t->off_b1),
                                      (const float*)(t->blob + t->off_w2), (const float*)(t->blob + t->off_b2), policy, value, spx::EngineDev{}, 1, -1LL));
    } else {
        const int grid = (int)(groups < t->sm_count ? groups : t->sm_count);
        tower_kernel<1><<<grid, NUM_THREADS, sizeof(SmemT<1>), st>>>((const unsigned long long*)own, (const unsigned long long*)opp, needs_eval, n,
                                                            t->n_layers, t->blob, (const float*)(t->blob + t->off_bias), t->head_buf,
                                                            0, nullptr, nullptr, nullptr, nullptr, nullptr, policy, value, spx::EngineDev{}, 1, -1LL);
    }
    spx::count_launch();
    SPX_CUDA_T(cudaGetLastError());
    if (e1) SPX_CUDA_T(cudaEventRecord(e1, st));
    if (t->fused) {   // the FC heads ran inside the tower kernel
        if (e2) SPX_CUDA_T(cudaEventRecord(e2, st));
        return 0;
    }
    auto hk = t->f16 ? heads_kernel<true> : heads_kernel<false>;
    hk<<<dim3((unsigned)((n + HB - 1) / HB), 2), 256, HEADS_SMEM, st>>>(
        t->head_buf, needs_eval, n, t->A, (const float*)(t->blob + t->off_polw), (const float*)(t->blob + t->off_polb),
        (const __nv_bfloat16*)(t->blob + t->off_w1t), (const float*)(t->blob + t->off_b1), (const float*)(t->blob + t->off_w2),
        (const float*)(t->blob + t->off_b2), policy, value);
    spx::count_launch();
    SPX_CUDA_T(cudaGetLastError());
    if (e2) SPX_CUDA_T(cudaEventRecord(e2, st));
    return 0;
}

#ifdef SPX_DBG_TRACE
int spx_debug_tick_trace(long long* host_out) { return (int)cudaMemcpyFromSymbol(host_out, spx::tower::g_tt, sizeof(long long) * 64 * 12); }
int spx_debug_trace(long long* host_out) { return (int)cudaMemcpyFromSymbol(host_out, spx::tower::g_trace, sizeof(long long) * (64 * 16 + 3 * 160)); }
#endif

/* n_ticks ticks (spx_advance + network evaluation each) in ONE launch: see tower_kernel<.., ENGINE = true>.
   balanced = 0: every game gets exactly n_ticks ticks (the launch lasts as long as its slowest SM pair);
   balanced = 1: work-conserving -- the launch holds n_ticks x (units of 14 games) network passes and every SM pair draws its next
   tick from that budget when it gets there, so pairs whose games search longer run fewer ticks and nobody waits. */
static int tick_fused_impl(spx_engine* e, spx_tower* t, int32_t n_ticks, float* policy, float* value, void* stream, int balanced) {
    if (!e || !t || !policy || !value) return spx::set_err(SPX_E_ARG, "spx_tick_fused: null argument%s", "");
    if (n_ticks <= 0) return 0;
    if (t->ncta != 2 || !t->fused) return spx::set_err(SPX_E_STATE, "spx_tick_fused: needs the SM-pair tower with fused heads%s", "");
    if (e->d.cfg.game != t->game || e->d.cfg.two_nets) return spx::set_err(SPX_E_ARG, "spx_tick_fused: one network, same game as the engine%s", "");
    if (e->d.K > 1) return spx::set_err(SPX_E_STATE, "spx_tick_fused: engines with search_threads > 1 tick with spx_advance + the network forward%s", "");
    const long long n = e->d.cfg.n_games, groups = (n + NB - 1) / NB, pairs = (groups + 1) / 2, max_pairs = t->sm_count / 2;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(2 * (pairs < max_pairs ? pairs : max_pairs)));
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = sizeof(SmemT<2>);
    cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    spx::EngineDev d = e->d;
    d.ecache = e->ecache;   // the evaluation cache (null = off) answers repeated requests inside the fused kernel only
    d.ecache_log2 = (unsigned)e->d.cfg.eval_cache_log2;
    d.ecache_tag[0] = (t->version << 8) | 1u;
    d.ecache_tag[1] = 0u;
    long long pass_budget = -1;
    if (balanced && n_ticks > 1) {   // the budget beyond every cluster's first tick; the ticket counter lives next to the tick counter
        pass_budget = (long long)(n_ticks - 1) * pairs;
        SPX_CUDA_T(cudaMemsetAsync(e->d.ticks + 1, 0, sizeof(unsigned long long), (cudaStream_t)stream));
    }
    auto kern = t->f16 ? (t->game == SPX_GAME_TICTACTOE ? tower_kernel<2, SPX_GAME_TICTACTOE, true, true> : tower_kernel<2, SPX_GAME_CONNECT4, true, true>)
                       : (t->game == SPX_GAME_TICTACTOE ? tower_kernel<2, SPX_GAME_TICTACTOE, true> : tower_kernel<2, SPX_GAME_CONNECT4, true>);
    SPX_CUDA_T(cudaLaunchKernelEx(&cfg, kern, (const unsigned long long*)e->d.leaf_own, (const unsigned long long*)e->d.leaf_opp,
                                  (const unsigned char*)e->d.needs_eval, n, t->n_layers, (const unsigned char*)t->blob,
                                  (const float*)(t->blob + t->off_bias), t->head_buf, 1, (const float*)(t->blob + t->off_polw),
                                  (const float*)(t->blob + t->off_polb), (const float*)(t->blob + t->off_b1), (const float*)(t->blob + t->off_w2),
                                  (const float*)(t->blob + t->off_b2), policy, value, d, (int)n_ticks, pass_budget));
    spx::count_launch();
    SPX_CUDA_T(cudaGetLastError());
    return 0;
}

int spx_tick_fused(spx_engine* e, spx_tower* t, int32_t n_ticks, float* policy, float* value, void* stream) {
    return tick_fused_impl(e, t, n_ticks, policy, value, stream, 0);
}
int spx_tick_fused_balanced(spx_engine* e, spx_tower* t, int32_t n_ticks, float* policy, float* value, void* stream) {
    return tick_fused_impl(e, t, n_ticks, policy, value, stream, 1);
}

int spx_tower_forward(spx_tower* t, const uint64_t* own, const uint64_t* opp, const uint8_t* needs_eval, int64_t n,
                      float* policy, float* value, void* stream) {
    return tower_forward_impl(t, own, opp, needs_eval, n, policy, value, (cudaStream_t)stream, nullptr, nullptr, nullptr);
}

/* same, recording caller-owned cudaEvent_t handles before the conv tower, between tower and heads, and after */
int spx_tower_forward_timed(spx_tower* t, const uint64_t* own, const uint64_t* opp, const uint8_t* needs_eval, int64_t n,
                            float* policy, float* value, void* stream, void* ev_start, void* ev_tower_done, void* ev_end) {
    return tower_forward_impl(t, own, opp, needs_eval, n, policy, value, (cudaStream_t)stream, (cudaEvent_t)ev_start,
                              (cudaEvent_t)ev_tower_done, (cudaEvent_t)ev_end);
}

}  // extern "C"

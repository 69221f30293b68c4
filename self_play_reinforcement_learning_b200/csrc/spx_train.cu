// spx_train.cu -- one SGD step of the ResidualTower on the device, hand-written for sm_100a.
//
// Replaces MCTreeSearch.loss + update_from_memory (games/algos/mcts.py:234-270) as UpdateWorker.update drives them
// (games/algos/updateworker.py:141-149): network.train() forward (BatchNorm on batch statistics, Dropout(0.5) on the two head
// activations; general/modules.py:88-107), loss = MSE(value, target) + (-sum(log p * tree_probs) / B), backward, and
// SGD(momentum, weight decay) on every parameter (self_play_parallel.py:193).
//
// Design (B200-first; a batch of 128 boards is 5376 board cells, so every layer is a small GEMM and the step is latency bound:
// ~340 short launches on one stream, no host round trip):
//   * Activations and gradients live in HBM as fp32 "planes": [C/4 chunks][rows][4 channels]; a row is a padded board cell
//     (56 rows per board, index col*7+row, zero guard cells -- the layout of spx_tower.cu), so a 3x3 tap is a constant row
//     shift and a 128-row tile of any tensor is C/4 contiguous 2 KB pieces that cp.async.bulk (TMA) moves as they lie.
//   * conv forward and backward-data: conv_tf32_kernel -- tcgen05.mma.cta_group::1.kind::tf32, M = 128 rows x N = 128 (64)
//     channels x K = 8, A = the activation tile read in place through shifted K-major no-swizzle descriptors (no im2col),
//     B = the weight slice of the tap streamed through a 4 x 32 KB TMA ring, D = fp32 in TMEM.  The epilogue stages the tile
//     in shared memory, adds the bias / the skip gradient, zeroes padding rows, writes planes, and produces the per-tile
//     BatchNorm sums.  Backward-data is the same kernel on the gradient planes with flipped taps and transposed weights.
//   * backward-weights: wgrad_kernel -- D[ci][co] += X^T dY over a chunk of rows: BOTH operands are read MN-major from planes
//     of the same shape (8 rows x 16 B = one core matrix with the row as the K index), so the tap shift is again an address
//     offset; three taps per CTA (3 x 128 TMEM columns), split-K over row chunks, fp32 partials reduced deterministically.
//     tcgen05 reads TF32 operands MN-major only from the 128-byte-swizzled layout, so this GEMM runs in kind::f16 on bf16
//     copies of the activations / conv-output gradients ([C/8 chunks][rows][8], written by the BatchNorm kernels); its K
//     dimension is the 5376 board cells of the batch, so the rounding errors of the products average out.
//   * BatchNorm (training mode) forward/backward, heads (Dropout, Linear layers, softmax, tanh), loss and SGD are fp32 CUDA-core
//     kernels; per-channel reductions are two-stage (per-tile partials, fp64 finalisation) so results do not depend on timing.
//   * TF32 (10-bit mantissa, fp32 range -- what stock PyTorch uses for fp32 convolutions on this GPU) for the GEMMs, fp32
//     everywhere else, fp32 master weights.  The reference trains under fp16 autocast without a GradScaler.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <new>
#include <vector>

#include "../../include/spx.h"

namespace spx {
int set_err(int code, const char* fmt, const char* detail);
void count_launch();
}  // namespace spx

#define SPX_CUDA_R(expr)                                                                  \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) return spx::set_err(SPX_E_CUDA, #expr ": %s", cudaGetErrorString(_e)); \
    } while (0)

namespace spx {
namespace train {

typedef unsigned long long ull;
constexpr int CH = 128, HEAD = 64, CELLS = 42, BOARD_ROWS = 56, GUARD = 8, TILE = 128, AROWS = TILE + 2 * GUARD;
constexpr int FLAT = 32 * CELLS, HID = 256, NA = 7;
constexpr int PIECE = AROWS * 16;            // bytes of one 4-channel chunk of an activation tile (144 rows x 16 B)
constexpr int RING_STAGES = 8, STAGE_MAX = 16384, NC = 64;   // weight ring; output channels per CTA of the conv kernel
constexpr float BN_EPS = 1e-5f, BN_MOM = 0.1f;

// ------------------------------------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, unsigned bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// the same copy delivered to the same shared-memory offset (and mbarrier) of every CTA of the cluster named in `mask`
__device__ __forceinline__ void tma_bulk_g2s_mc(void* dst, const void* src, unsigned bytes, void* bar, unsigned short mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ unsigned cluster_ctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned cluster_nctarank() { unsigned r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(void* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// completion of this thread's MMAs -> the same mbarrier in every CTA of the cluster named in `mask`
__device__ __forceinline__ void tc_commit_mc(void* bar, unsigned short mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(unsigned d_tmem, ull adesc, ull bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(unsigned d_tmem, ull adesc, ull bdesc, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// the same MMAs with the descriptors passed as (low word, constant high word): the issuing thread advances a descriptor with ONE
// 32-bit add (low word = LBO >> 4 << 16 | address >> 4; high word = SBO >> 4 | version bit) -- the issue loop is a single
// thread's instruction stream and paces the kernel otherwise (measured: 129 cycles per MMA with make_desc in the loop)
__device__ __forceinline__ void tc_mma_tf32_lo(unsigned d_tmem, unsigned a_lo, unsigned a_hi, unsigned b_lo, unsigned b_hi, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tc_mma_f16_lo(unsigned d_tmem, unsigned a_lo, unsigned a_hi, unsigned b_lo, unsigned b_hi, unsigned idesc, unsigned accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ unsigned desc_lo(unsigned saddr, unsigned lbo) { return ((saddr >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ unsigned desc_hi(unsigned sbo) { return ((sbo >> 4) & 0x3FFFu) | (1u << 14); }
__device__ __forceinline__ void tc_ld32(unsigned taddr, float (&v)[32]) {
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// shared-memory matrix descriptor, no swizzle: start address, leading / stride byte offsets (all >> 4), descriptor version 1.
//   K-major  operand: core matrix = 8 rows x 16 B; SBO = stride between 8-row groups, LBO = stride between the 16-byte K chunks.
//   MN-major operand: 16 B = 4 consecutive M (N) elements; SBO = stride between such groups, 8 consecutive K steps are 16 B
//   apart, LBO = stride between groups of 8 K steps (cute::UMMA make_umma_desc, canonical INTERLEAVE layouts).
__device__ __forceinline__ ull make_desc(unsigned saddr, unsigned lbo, unsigned sbo) {
    return (ull)((saddr >> 4) & 0x3FFFu) | ((ull)((lbo >> 4) & 0x3FFFu) << 16) | ((ull)((sbo >> 4) & 0x3FFFu) << 32) | (1ULL << 46);
}
// instruction descriptor, kind::tf32: D = f32 (bit 4), A/B format TF32 = 2 (bits 7 / 10), a_major / b_major (bits 15 / 16: 1 = MN-major),
// N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr unsigned make_idesc_tf32(int M, int N, bool mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (mn_major ? ((1u << 15) | (1u << 16)) : 0u) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

// kind::f16 with bf16 operands (format 1), both MN-major
__host__ __device__ constexpr unsigned make_idesc_bf16_mn(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((unsigned)(N >> 3) << 17) | ((unsigned)(M >> 4) << 24);
}

// Programmatic dependent launch: every kernel of the step is launched with programmaticStreamSerialization, so its CTAs are
// scheduled while the previous kernel of the stream is still draining; pdl_sync() lets the NEXT kernel start the same way and then
// waits until the previous grid has completed and its memory is visible.  (A step is ~340 kernels of ~10 us: the launch gaps were
// ~15 % of it.)  Must run before the first access to anything an earlier kernel wrote.
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ geometry
// global row g (0 .. Rg-1) of a plane tensor: 8 guard rows, then 56 rows per board (cell (col, row) at col*7 + row), 8 guard rows
__device__ __forceinline__ bool row_real(int g, int B, int& b, int& cell) {
    const int r = g - GUARD;
    if (r < 0) { b = 0; cell = 0; return false; }
    b = r / BOARD_ROWS;
    const int p = r - b * BOARD_ROWS, col = p / 7, row = p - col * 7;
    cell = col * 6 + row;
    return b < B && p < 49 && row < 6;
}
__device__ __forceinline__ int tap_shift(int taps, int tap) { return taps == 9 ? (tap / 3 - 1) * 7 + (tap % 3 - 1) : 0; }

// a per-channel vector that may live in two tensors (the fused policy|value head conv: channels [0, split) and [split, N))
struct Seg2 {
    float* p0; float* p1; int split;
    __device__ __forceinline__ float* at(int c) const { return c < split ? p0 + c : p1 + (c - split); }
};

// ------------------------------------------------------------------------------------------------ conv forward / backward-data
struct ConvP {
    const float* in; int in_chunks;     // planes of the input tensor, K / 4
    const float* w; int taps, ks_per_tap;   // packed weights [N / 64 halves][tap][kstep][2 k-chunks][64][4], K / 8 K-steps per tap
    int N;                              // output channels (64 or 128)
    float* out;                         // planes, N / 4 chunks
    Seg2 bias; int has_bias;
    const float* add;                   // optional planes added on real rows (skip-connection gradient)
    float* stat;                        // optional [tiles][2][N]: per-tile sum / sum of squares over real rows
    int B, Rg;
    long long* trace;                   // debug: clock64 at 8 points of CTA 0 (null in production)
};
#define SPX_CT(k) do { if (p.trace && blockIdx.x == 0 && blockIdx.y == 0) p.trace[k] = clock64(); } while (0)
constexpr int CONV_CLUSTER = 1;   // SPX_CONV_CLUSTER=7: seven row tiles share one multicast weight stream (measured slower, see below)
constexpr int CONV_SMEM = (CH / 4) * PIECE + RING_STAGES * STAGE_MAX + 256 + TILE * 4;

__global__ void __launch_bounds__(256, 1) conv_tf32_kernel(const ConvP p) {
    pdl_sync();
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* atile = smem;
    unsigned char* ring = smem + (CH / 4) * PIECE;
    ull* bars = reinterpret_cast<ull*>(ring + RING_STAGES * STAGE_MAX);   // full[8] empty[8] abar accbar
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 20);
    int* realrow = reinterpret_cast<int*>(ring + RING_STAGES * STAGE_MAX + 256);
    ull* full = bars; ull* empty = bars + RING_STAGES; ull* abar = bars + 2 * RING_STAGES; ull* accbar = abar + 1;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, r0 = tile * TILE;      // tile row m <-> global row r0 + GUARD + m; A-tile row a <-> global row r0 + a
    const int N = p.N, n0 = blockIdx.y * NC;            // this CTA computes output channels [n0, n0 + 64)
    const int ks_stage = p.ks_per_tap < 8 ? p.ks_per_tap : 8, stages_per_tap = p.ks_per_tap / ks_stage, total = p.taps * stages_per_tap;
    const unsigned stage_bytes = (unsigned)(ks_stage * 2 * NC * 16);
    // Optional (SPX_CONV_CLUSTER=7): the CTAs of a cluster work on consecutive row tiles of the SAME 64 output channels, i.e. they
    // consume the same weight stream: stage i is fetched once, by CTA i % CL, and multicast into every CTA's ring slot; a slot is
    // refilled when the MMAs of ALL CTAs have released it.  Measured: the kernel's own time does not change (it is not bound by
    // the 33 MB of L2 -> SM weight reads per layer) and the step gets SLOWER, 3.82 -> 5.05 ms: a 7-CTA cluster needs 7 free SMs of
    // one GPC at once, which serialises the tails of the ~340 short dependent kernels.  Default: no clusters (CL = 1).
    const unsigned crank = cluster_ctarank(), CL = cluster_nctarank();
    const unsigned short cmask = (unsigned short)((1u << CL) - 1u);
    if (tid == 0) SPX_CT(0);

    if (tid == 0) {
        for (int i = 0; i < RING_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], CL); }
        mbar_init(abar, 1); mbar_init(accbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid < TILE) { int b, cell; realrow[tid] = row_real(r0 + GUARD + tid, p.B, b, cell) ? 1 : 0; }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();      // every CTA's barriers are initialised before any multicast copy / commit can reach them
    tc_fence_after();
    const unsigned tmem = *tmem_slot;
    if (tid == 0) SPX_CT(1);

    if (warp == 0 && lane == 0) {
        // ---- TMA producer: the activation tile (one 2304-byte piece per 4-channel chunk), then the weight stages in MMA order
        mbar_expect_tx(abar, (unsigned)(p.in_chunks * PIECE));
        for (int c = 0; c < p.in_chunks; ++c)
            tma_bulk_g2s(atile + c * PIECE, p.in + ((size_t)c * p.Rg + r0) * 4, PIECE, abar);
        const unsigned char* wsrc = reinterpret_cast<const unsigned char*>(p.w) + (size_t)blockIdx.y * total * stage_bytes;   // this half's slices
        for (int i = 0; i < total; ++i) {
            const int slot = i % RING_STAGES;
            if (i >= RING_STAGES) mbar_wait(&empty[slot], (unsigned)((i / RING_STAGES - 1) & 1));   // released by all CL CTAs
            mbar_expect_tx(&full[slot], stage_bytes);                                                // every CTA arms its own barrier
            if ((unsigned)i % CL == crank)
                tma_bulk_g2s_mc(ring + slot * STAGE_MAX, wsrc + (size_t)i * stage_bytes, stage_bytes, &full[slot], cmask);
        }
        SPX_CT(2);
    } else if (warp == 1 && lane == 0) {
        // ---- MMA issuer
        // consecutive MMAs go to FOUR accumulators in turn (K split four ways, summed by the epilogue): back-to-back MMAs into one
        // accumulator wait for each other (~150 instead of ~35 cycles each), and every layer issues >= 8 of them
        const unsigned idesc = make_idesc_tf32(TILE, NC, false);
        const unsigned hi = desc_hi(128);                                                     // SBO = 128 B for both operands
        const unsigned a_lo0 = desc_lo(smem_u32(atile) + GUARD * 16, PIECE);                  // tile row 0, K-step 0, no shift
        const unsigned b_lo0 = desc_lo(smem_u32(ring), NC * 16);
        constexpr unsigned A_KSTEP = (2 * PIECE) >> 4, B_KSTEP = (2 * NC * 16) >> 4, B_SLOT = STAGE_MAX >> 4;
        unsigned mm = 0;
        mbar_wait(abar, 0);
        SPX_CT(3);
        for (int i = 0; i < total; ++i) {
            const int slot = i % RING_STAGES;
            mbar_wait(&full[slot], (unsigned)((i / RING_STAGES) & 1));
            tc_fence_after();
            const int tap = i / stages_per_tap, half = i - tap * stages_per_tap;
            unsigned a_lo = a_lo0 + (unsigned)tap_shift(p.taps, tap) + (unsigned)(half * ks_stage) * A_KSTEP;   // a row is one 16-byte address unit
            unsigned b_lo = b_lo0 + (unsigned)slot * B_SLOT;
            if (ks_stage == 8) {
#pragma unroll
                for (int ks = 0; ks < 8; ++ks) {
                    tc_mma_tf32_lo(tmem + (mm & 3u) * NC, a_lo, hi, b_lo, hi, idesc, mm >= 4u ? 1u : 0u);
                    ++mm; a_lo += A_KSTEP; b_lo += B_KSTEP;
                }
            } else {
                for (int ks = 0; ks < ks_stage; ++ks) {
                    tc_mma_tf32_lo(tmem + (mm & 3u) * NC, a_lo, hi, b_lo, hi, idesc, mm >= 4u ? 1u : 0u);
                    ++mm; a_lo += A_KSTEP; b_lo += B_KSTEP;
                }
            }
            tc_commit_mc(&empty[slot], cmask);
        }
        tc_commit(accbar);
        SPX_CT(4);
    } else if (warp >= 4) {
        // ---- epilogue: TMEM -> shared staging [128 rows][64 + 1] (reuses the weight ring once every MMA has retired)
        const int et = tid - 128;                 // TMEM lane == tile row
        float* st = reinterpret_cast<float*>(ring);
        constexpr int ld = NC + 1;
        mbar_wait(accbar, 0);
        tc_fence_after();
        if (et == 0) SPX_CT(5);
#pragma unroll
        for (int cg = 0; cg < NC / 32; ++cg) {
            float v[32], u[32];
            const unsigned tl = tmem + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)(cg * 32);
            tc_ld32(tl, v);
#pragma unroll
            for (int a = 1; a < 4; ++a) {          // the four K-split accumulators, summed in a fixed order
                tc_ld32(tl + (unsigned)(a * NC), u);
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] += u[j];
            }
#pragma unroll
            for (int j = 0; j < 32; ++j) st[et * ld + cg * 32 + j] = v[j];
        }
        tc_fence_before();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (et == 0) SPX_CT(6);
        // pass 1, two threads per channel (64 rows each): bias, padding rows -> 0, per-tile BatchNorm sums over the real rows
        {
            const int c = et & (NC - 1), hrow = et >> 6;
            const float bv = p.has_bias ? *p.bias.at(n0 + c) : 0.f;
            float s = 0.f, sq = 0.f;
#pragma unroll 8
            for (int m = hrow * 64; m < hrow * 64 + 64; ++m) {
                float v = st[m * ld + c] + bv;
                if (!realrow[m]) v = 0.f;
                s += v; sq = fmaf(v, v, sq);          // padding rows contribute exact zeros
                st[m * ld + c] = v;
            }
            float* red = st + TILE * ld;          // [2 row halves][2 stats][64]
            red[(hrow * 2) * NC + c] = s; red[(hrow * 2 + 1) * NC + c] = sq;
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (p.stat && et < NC) {
            const float* red = st + TILE * ld;
            p.stat[((size_t)tile * 2) * N + n0 + et] = red[et] + red[2 * NC + et];
            p.stat[((size_t)tile * 2 + 1) * N + n0 + et] = red[NC + et] + red[3 * NC + et];
        }
        // pass 2, thread = row: planes out (16 B per row and chunk, coalesced over rows); the optional skip-gradient loads of
        // four chunks are issued together before they are used (one L2 round trip per four chunks instead of one each)
        {
            const int m = et;     // 128 threads = 128 rows; 16 chunks each
            const bool addr = p.add && realrow[m];
            const size_t o0 = (size_t)(n0 / 4) * p.Rg + r0 + GUARD + m;
#pragma unroll
            for (int c4 = 0; c4 < NC / 4; c4 += 4) {
                float4 a[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) a[k] = addr ? reinterpret_cast<const float4*>(p.add)[o0 + (size_t)(c4 + k) * p.Rg] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ch = c4 + k;
                    const float4 v = make_float4(st[m * ld + 4 * ch] + a[k].x, st[m * ld + 4 * ch + 1] + a[k].y, st[m * ld + 4 * ch + 2] + a[k].z, st[m * ld + 4 * ch + 3] + a[k].w);
                    reinterpret_cast<float4*>(p.out)[o0 + (size_t)ch * p.Rg] = v;
                }
            }
        }
        if (et == 0) SPX_CT(7);
    }
    tc_fence_before();
    __syncthreads();
    if (CL > 1) cluster_sync_all();      // no CTA leaves while another one may still signal its barriers
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

// ------------------------------------------------------------------------------------------------ backward-weights
struct WgP {
    const __nv_bfloat16* x;   // bf16 planes [16 chunks][rows][8] of the layer's input activations (128 channels)
    const __nv_bfloat16* dy;  // bf16 planes [N/8 chunks][rows][8] of the gradient w.r.t. the conv output, zero on padding rows
    int N, taps;         // taps 9 -> three CTAs (blockIdx.y) of three taps each; taps 1 -> one
    float* partial;      // [S][taps][128 ci][N co]
    int Rg, nchunks;     // row chunks of 128 rows
    int a_lbo, a_sbo, b_lbo, b_sbo;   // descriptor strides (bytes)
};
constexpr int WG_SMEM = (CH / 8) * PIECE + (CH / 8) * TILE * 16 + 256;

__global__ void __launch_bounds__(256, 1) wgrad_kernel(const WgP p) {
    pdl_sync();
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* xt = smem;                               // 16 chunks x 144 rows x 16 B
    unsigned char* dt = smem + (CH / 8) * PIECE;            // N/8 chunks x 128 rows x 16 B
    ull* bars = reinterpret_cast<ull*>(dt + (CH / 8) * TILE * 16);   // ldbar, mmabar, accbar
    unsigned* tmem_slot = reinterpret_cast<unsigned*>(bars + 4);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int N = p.N, S = gridDim.x, s = blockIdx.x;
    const int ntap = p.taps == 9 ? 3 : 1, tap0 = blockIdx.y * 3, nsub = ntap == 1 ? 4 : 1;   // accumulators per tap
    if (tid == 0) {
        mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); mbar_init(&bars[2], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const unsigned tmem = *tmem_slot;
    if (warp == 0 && lane == 0) {
        const unsigned idesc = make_idesc_bf16_mn(CH, N);
        const unsigned bytes = (unsigned)((CH / 8) * PIECE + (N / 8) * TILE * 16);
        int it = 0;
        for (int j = s; j < p.nchunks; j += S, ++it) {
            if (it > 0) mbar_wait(&bars[1], (unsigned)((it - 1) & 1));     // the previous chunk's MMAs have read the tiles
            mbar_expect_tx(&bars[0], bytes);
            for (int c = 0; c < CH / 8; ++c) tma_bulk_g2s(xt + c * PIECE, p.x + ((size_t)c * p.Rg + (size_t)j * TILE) * 8, PIECE, &bars[0]);
            for (int c = 0; c < N / 8; ++c) tma_bulk_g2s(dt + c * TILE * 16, p.dy + ((size_t)c * p.Rg + (size_t)j * TILE + GUARD) * 8, TILE * 16, &bars[0]);
            mbar_wait(&bars[0], (unsigned)(it & 1));
            tc_fence_after();
            // consecutive MMAs alternate between independent accumulators (back-to-back MMAs into one accumulator wait for each
            // other): the three taps of the group, or -- one tap only (1x1 head conv) -- four K-split accumulators summed by the epilogue
            {
                const unsigned a_hi = desc_hi((unsigned)p.a_sbo), b_hi = desc_hi((unsigned)p.b_sbo);
                const unsigned a_lo0 = desc_lo(smem_u32(xt) + GUARD * 16, (unsigned)p.a_lbo), b_lo0 = desc_lo(smem_u32(dt), (unsigned)p.b_lbo);
                unsigned sh[3];
                for (int t = 0; t < ntap; ++t) sh[t] = (unsigned)tap_shift(p.taps, tap0 + t);
                // A[ci][k = row] and B[co][k = row], both MN-major: 8 channels per 16 B, chunk stride = SBO, rows 16 B apart, the two
                // groups of 8 rows of one K = 16 step LBO = 128 B apart; a K step advances both start addresses by 16 rows
#pragma unroll
                for (int kk = 0; kk < TILE / 16; ++kk) {
                    for (int t = 0; t < ntap; ++t) {
                        const int sub = ntap == 1 ? (kk & 3) : 0;
                        const bool first = it == 0 && (ntap == 1 ? kk < 4 : kk == 0);
                        tc_mma_f16_lo(tmem + (unsigned)((t * nsub + sub) * N), a_lo0 + sh[t] + (unsigned)(16 * kk), a_hi, b_lo0 + (unsigned)(16 * kk), b_hi, idesc,
                                      first ? 0u : 1u);
                    }
                }
            }
            tc_commit(&bars[1]);
        }
        tc_commit(&bars[2]);
    } else if (warp >= 4) {
        const int ci = tid - 128;
        mbar_wait(&bars[2], 0);
        tc_fence_after();
        for (int t = 0; t < ntap; ++t) {
            float* dst = p.partial + (((size_t)s * p.taps + tap0 + t) * CH + ci) * N;
            for (int cg = 0; cg < N / 32; ++cg) {
                float v[32], u[32];
                const unsigned tl = tmem + ((unsigned)((warp & 3) * 32) << 16) + (unsigned)(t * nsub * N + cg * 32);
                tc_ld32(tl, v);
                for (int a = 1; a < nsub; ++a) {
                    tc_ld32(tl + (unsigned)(a * N), u);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] += u[j];
                }
#pragma unroll
                for (int j = 0; j < 32; j += 4) reinterpret_cast<float4*>(dst + cg * 32)[j >> 2] = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// grad[(co * CIN + ci) * taps + tap] = sum_s partial[((s * taps + tap) * CINP + ci) * N + co]   (PyTorch [co][ci][kh][kw]);
// channels >= split go to a second tensor (the value head conv)
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int S, int taps, int CINP, int CIN, int N, float* g0, float* g1, int split) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= taps * CIN * N) return;
    const int co = i % N, ci = (i / N) % CIN, tap = i / (N * CIN);
    float acc = 0.f;
    for (int s = 0; s < S; ++s) acc += partial[(((size_t)s * taps + tap) * CINP + ci) * N + co];
    float* g = co < split ? g0 + ((size_t)co * CIN + ci) * taps + tap : g1 + ((size_t)(co - split) * CIN + ci) * taps + tap;
    *g = acc;
}

// stem backward-weights on the CUDA cores (3 input planes): partial[s][tap][ci < 3][co]
__global__ void stem_wgrad_kernel(const float* __restrict__ x0, const float* __restrict__ dy, float* partial, int Rg, int R, int rows_per_split) {
    pdl_sync();
    const int tap = blockIdx.y, s = blockIdx.x, co = threadIdx.x;
    const int shift = tap_shift(9, tap);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    const int g_end = min(GUARD + (s + 1) * rows_per_split, GUARD + R);
    for (int g = GUARD + s * rows_per_split; g < g_end; ++g) {
        const float d = dy[((size_t)(co >> 2) * Rg + g) * 4 + (co & 3)];
        const float4 x = reinterpret_cast<const float4*>(x0)[g + shift];     // chunk 0 of the input planes: (empty, own, enemy, 0)
        a0 = fmaf(x.x, d, a0); a1 = fmaf(x.y, d, a1); a2 = fmaf(x.z, d, a2);
    }
    float* dst = partial + ((size_t)s * 9 + tap) * 3 * CH;
    dst[co] = a0; dst[CH + co] = a1; dst[2 * CH + co] = a2;
}

// ------------------------------------------------------------------------------------------------ BatchNorm (training mode)
struct BnP {
    const float* y;        // conv output planes (bias included)
    float* a;              // activation planes out (forward) / activation planes in (backward: ReLU mask)
    const float* res;      // forward: optional residual planes
    const float* stat;     // forward: [tiles][2][N] from the conv epilogue
    Seg2 gamma, beta, rmean, rvar;
    float* mean; float* invstd;   // [N] saved for the backward pass
    const float* g;        // backward: gradient planes w.r.t. the activation
    float* dy;             // backward: gradient planes w.r.t. the conv output
    float* skip;           // backward: optional copy of dz (gradient through the identity branch)
    float* part;           // backward: [tiles][2][N] partial sums (sum dz, sum dz * xhat)
    Seg2 dgamma, dbeta, dbias;
    __nv_bfloat16* a16;    // forward: optional bf16 copy of the activation planes [N/8][rows][8] (operand of the backward-weights GEMM)
    __nv_bfloat16* dy16;   // backward: bf16 copy of dy, same layout
    int N, B, Rg, tiles, n_real;
};
// 4 consecutive channels of fp32 chunk `ch` -> their 8 bytes inside the 16-byte unit of bf16 chunk ch / 2
__device__ __forceinline__ void store_bf16x4(__nv_bfloat16* planes, int ch, int Rg, int g, float4 v) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
    uint2 w;
    w.x = *reinterpret_cast<const unsigned*>(&lo); w.y = *reinterpret_cast<const unsigned*>(&hi);
    *reinterpret_cast<uint2*>(planes + ((size_t)(ch >> 1) * Rg + g) * 8 + (ch & 1) * 4) = w;
}

// The three BatchNorm kernels run on a grid of (row tiles, N / 16) blocks: a block owns 128 rows x 16 channels (4 plane chunks;
// thread = row x chunk pair), so that a 5376-row tensor is spread over a few hundred blocks instead of 56.
constexpr int BN_CG = 16;    // channels per block

// per-channel totals of the per-tile partial sums part[tile][2][N] for the block's 16 channels, in fp64 and in a fixed order
// (8 tile subsets per (stat, channel) pair, then the subsets in order): tot[0][c] / tot[1][c]
__device__ __forceinline__ void bn_totals(const float* __restrict__ part, int tiles, int N, int c0, double (*tot)[BN_CG]) {
    __shared__ double sub[8][2 * BN_CG];
    const int tid = threadIdx.x, cs = tid & 31, subset = tid >> 5, stat = cs >> 4, c = cs & 15;
    double acc = 0.0;
    for (int t = subset; t < tiles; t += 8) acc += (double)part[((size_t)t * 2 + stat) * N + c0 + c];
    sub[subset][cs] = acc;
    __syncthreads();
    if (tid < 2 * BN_CG) {
        double a2 = 0.0;
        for (int k = 0; k < 8; ++k) a2 += sub[k][tid];
        tot[tid >> 4][tid & 15] = a2;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(256) bn_fwd_kernel(const BnP p) {
    pdl_sync();
    __shared__ double tot[2][BN_CG];
    __shared__ float sc[BN_CG], sh[BN_CG];
    const int tid = threadIdx.x, N = p.N, r0 = blockIdx.x * TILE, c0 = blockIdx.y * BN_CG;
    bn_totals(p.stat, p.tiles, N, c0, tot);
    if (tid < BN_CG) {
        const int c = c0 + tid;
        const double mean = tot[0][tid] / p.n_real, var = fmax(tot[1][tid] / p.n_real - mean * mean, 0.0);
        const float invstd = (float)(1.0 / sqrt(var + (double)BN_EPS));
        const float gm = *p.gamma.at(c), bt = *p.beta.at(c);
        sc[tid] = gm * invstd;
        sh[tid] = bt - (float)mean * gm * invstd;
        if (blockIdx.x == 0) {
            p.mean[c] = (float)mean; p.invstd[c] = invstd;
            float* rm = p.rmean.at(c); float* rv = p.rvar.at(c);     // nn.BatchNorm2d: momentum 0.1, unbiased variance in the running estimate
            *rm = (1.f - BN_MOM) * *rm + BN_MOM * (float)mean;
            *rv = (1.f - BN_MOM) * *rv + BN_MOM * (float)(var * (double)p.n_real / (double)(p.n_real - 1));
        }
    }
    __syncthreads();
    const int m = tid & (TILE - 1), g = r0 + GUARD + m;
    int b, cell;
    const bool real = row_real(g, p.B, b, cell);
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) {
        const int lc = (tid >> 7) + 2 * k2, ch = (c0 >> 2) + lc;      // local / global plane chunk
        const size_t o = (size_t)ch * p.Rg + g;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (real) {
            const float4 y = reinterpret_cast<const float4*>(p.y)[o];
            v.x = fmaf(y.x, sc[4 * lc], sh[4 * lc]); v.y = fmaf(y.y, sc[4 * lc + 1], sh[4 * lc + 1]);
            v.z = fmaf(y.z, sc[4 * lc + 2], sh[4 * lc + 2]); v.w = fmaf(y.w, sc[4 * lc + 3], sh[4 * lc + 3]);
            if (p.res) { const float4 r = reinterpret_cast<const float4*>(p.res)[o]; v.x += r.x; v.y += r.y; v.z += r.z; v.w += r.w; }
            v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
        }
        reinterpret_cast<float4*>(p.a)[o] = v;
        if (p.a16) store_bf16x4(p.a16, ch, p.Rg, g, v);
    }
}

// dz = g * (a > 0); per-tile sums of dz and dz * xhat per channel (warp shuffles, then the block's 4 warps per chunk pair in order)
__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const BnP p) {
    pdl_sync();
    __shared__ float wsum[8][2][8];      // [warp][stat][channel within the warp's two chunks]
    const int tid = threadIdx.x, N = p.N, r0 = blockIdx.x * TILE, c0 = blockIdx.y * BN_CG, lane = tid & 31, warp = tid >> 5;
    const int m = tid & (TILE - 1), g = r0 + GUARD + m;
    int b, cell;
    const bool real = row_real(g, p.B, b, cell);
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) {
        const int lc = (tid >> 7) + 2 * k2, ch = (c0 >> 2) + lc;
        const size_t o = (size_t)ch * p.Rg + g;
        float d[4] = {0.f, 0.f, 0.f, 0.f}, dx[4] = {0.f, 0.f, 0.f, 0.f};
        if (real) {
            const float4 gg = reinterpret_cast<const float4*>(p.g)[o], a = reinterpret_cast<const float4*>(p.a)[o], y = reinterpret_cast<const float4*>(p.y)[o];
            const float gv[4] = {gg.x, gg.y, gg.z, gg.w}, av[4] = {a.x, a.y, a.z, a.w}, yv[4] = {y.x, y.y, y.z, y.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                d[k] = av[k] > 0.f ? gv[k] : 0.f;
                dx[k] = d[k] * ((yv[k] - p.mean[4 * ch + k]) * p.invstd[4 * ch + k]);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) { d[k] += __shfl_xor_sync(0xffffffffu, d[k], off); dx[k] += __shfl_xor_sync(0xffffffffu, dx[k], off); }
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 4; ++k) { wsum[warp][0][4 * k2 + k] = d[k]; wsum[warp][1][4 * k2 + k] = dx[k]; }
        }
    }
    __syncthreads();
    if (tid < 2 * BN_CG) {
        // channel lc4 = local channel 0..15 = chunk lc (0..3) * 4 + k; chunk lc belongs to thread half (lc & 1) and k2 = lc >> 1; warps 4*half .. 4*half + 3
        const int stat = tid >> 4, lc4 = tid & 15, lc = lc4 >> 2, k = lc4 & 3, half = lc & 1, k2 = lc >> 1;
        float acc = 0.f;
        for (int w = 0; w < 4; ++w) acc += wsum[4 * half + w][stat][4 * k2 + k];
        p.part[((size_t)blockIdx.x * 2 + stat) * N + c0 + lc4] = acc;
    }
}

// dy = gamma * invstd * (dz - mean(dz) - xhat * mean(dz * xhat));  dgamma = sum dz * xhat, dbeta = sum dz, dbias = 0 (exactly:
// a per-channel constant added before training-mode BatchNorm does not change its output)
__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const BnP p) {
    pdl_sync();
    __shared__ double tot[2][BN_CG];
    __shared__ float m1[BN_CG], m2[BN_CG], sc[BN_CG], mu[BN_CG], is[BN_CG];
    const int tid = threadIdx.x, N = p.N, r0 = blockIdx.x * TILE, c0 = blockIdx.y * BN_CG;
    bn_totals(p.part, p.tiles, N, c0, tot);
    if (tid < BN_CG) {
        const int c = c0 + tid;
        m1[tid] = (float)(tot[0][tid] / p.n_real); m2[tid] = (float)(tot[1][tid] / p.n_real);
        mu[tid] = p.mean[c]; is[tid] = p.invstd[c];
        sc[tid] = *p.gamma.at(c) * is[tid];
        if (blockIdx.x == 0) { *p.dgamma.at(c) = (float)tot[1][tid]; *p.dbeta.at(c) = (float)tot[0][tid]; *p.dbias.at(c) = 0.f; }
    }
    __syncthreads();
    const int m = tid & (TILE - 1), g = r0 + GUARD + m;
    int b, cell;
    const bool real = row_real(g, p.B, b, cell);
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) {
        const int lc = (tid >> 7) + 2 * k2, ch = (c0 >> 2) + lc;
        const size_t o = (size_t)ch * p.Rg + g;
        float4 out = make_float4(0.f, 0.f, 0.f, 0.f), dzv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (real) {
            const float4 gg = reinterpret_cast<const float4*>(p.g)[o], a = reinterpret_cast<const float4*>(p.a)[o], y = reinterpret_cast<const float4*>(p.y)[o];
            const float gv[4] = {gg.x, gg.y, gg.z, gg.w}, av[4] = {a.x, a.y, a.z, a.w}, yv[4] = {y.x, y.y, y.z, y.w};
            float r[4], dz[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int c = 4 * lc + k;
                dz[k] = av[k] > 0.f ? gv[k] : 0.f;
                const float xh = (yv[k] - mu[c]) * is[c];
                r[k] = sc[c] * (dz[k] - m1[c] - xh * m2[c]);
            }
            out = make_float4(r[0], r[1], r[2], r[3]);
            dzv = make_float4(dz[0], dz[1], dz[2], dz[3]);
        }
        reinterpret_cast<float4*>(p.dy)[o] = out;
        store_bf16x4(p.dy16, ch, p.Rg, g, out);
        if (p.skip) reinterpret_cast<float4*>(p.skip)[o] = dzv;
    }
}

// ------------------------------------------------------------------------------------------------ input planes
// planes [B][3][7][6] (preprocess, general/modules.py:115-125) -> X0: chunk 0 = (empty, own, enemy, 0), chunk 1 = 0
__global__ void input_kernel(const float* __restrict__ planes, float* x0, int B, int Rg) {
    pdl_sync();
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= Rg) return;
    int b, cell;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row_real(g, B, b, cell)) {
        const float* src = planes + (size_t)b * 3 * CELLS + cell;
        v = make_float4(src[0], src[CELLS], src[2 * CELLS], 0.f);
    }
    reinterpret_cast<float4*>(x0)[g] = v;
    reinterpret_cast<float4*>(x0)[(size_t)Rg + g] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// ------------------------------------------------------------------------------------------------ heads
__device__ __forceinline__ ull mix64(ull z) { z += 0x9E3779B97F4A7C15ULL; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL; z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL; return z ^ (z >> 31); }

struct HeadP {
    const float* ah;            // head activation planes (64 channels: 0..31 policy, 32..63 value), post BN + ReLU
    const unsigned char* mask_in;   // optional injected dropout keep-mask [B][2][FLAT]
    unsigned char* mask;        // keep-mask actually used [B][2][FLAT]
    float* xd;                  // activations after dropout [B][2][FLAT]
    const float* wp; const float* bp;          // linear_policy [7][1344], [7]
    const float* w1t; const float* b1;         // fc_value transposed [1344][256], [256]
    const float* w1;                           // fc_value [256][1344]
    const float* w2; const float* b2;          // linear_output [256], [1]
    const float* tree_probs; const float* target;
    float* probs; float* value; float* h1;     // [B][7], [B], [B][256]
    float* dlogit; float* du; float* dh1;      // gradients w.r.t. logits, tanh input, hidden layer (post-ReLU masked)
    float* lossb;                              // [B][2] per-board value / policy loss terms
    float* gah;                                // gradient planes w.r.t. ah (64 channels)
    ull seed, step;
    int B, Rg;
};

__global__ void __launch_bounds__(256) head_fwd_kernel(const HeadP p) {
    pdl_sync();
    __shared__ float xs[2][FLAT];
    __shared__ float logit[8], red[8];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int k = tid; k < 2 * FLAT; k += 256) {
        const int head = k / FLAT, kk = k - head * FLAT, c = kk / CELLS + 32 * head, cell = kk % CELLS;
        const int g = GUARD + b * BOARD_ROWS + (cell / 6) * 7 + cell % 6;
        const float x = p.ah[((size_t)(c >> 2) * p.Rg + g) * 4 + (c & 3)];
        const size_t o = ((size_t)b * 2 + head) * FLAT + kk;
        const unsigned char keep = p.mask_in ? p.mask_in[o] : (unsigned char)(mix64(mix64(p.seed ^ (p.step * 0x632BE59BD9B4E019ULL)) + o) >> 63);
        const float v = keep ? 2.f * x : 0.f;                         // nn.Dropout(p = 0.5), training mode
        xs[head][kk] = v; p.xd[o] = v; p.mask[o] = keep;
    }
    __syncthreads();
    if (warp < NA) {                                                  // linear_policy
        float acc = 0.f;
        for (int k = lane; k < FLAT; k += 32) acc = fmaf(xs[0][k], p.wp[warp * FLAT + k], acc);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) logit[warp] = acc + p.bp[warp];
    }
    float h = p.b1[tid];                                              // fc_value + ReLU, thread = hidden unit
    for (int k = 0; k < FLAT; ++k) h = fmaf(xs[1][k], p.w1t[(size_t)k * HID + tid], h);
    h = fmaxf(h, 0.f);
    p.h1[(size_t)b * HID + tid] = h;
    float part = h * p.w2[tid];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (tid == 0) {
        float u = p.b2[0];
        for (int w = 0; w < 8; ++w) u += red[w];
        const float v = tanhf(u), t = p.target[b];
        p.value[b] = v;
        const float invB = 1.f / (float)p.B;
        p.du[b] = 2.f * (v - t) * invB * (1.f - v * v);
        float mx = logit[0];
        for (int a = 1; a < NA; ++a) mx = fmaxf(mx, logit[a]);
        float z = 0.f;
        for (int a = 0; a < NA; ++a) z += expf(logit[a] - mx);
        const float lz = logf(z);
        float pi_sum = 0.f, lp = 0.f;
        for (int a = 0; a < NA; ++a) pi_sum += p.tree_probs[b * NA + a];
        for (int a = 0; a < NA; ++a) {
            const float logp = logit[a] - mx - lz, pr = expf(logp), pi = p.tree_probs[b * NA + a];
            p.probs[b * NA + a] = pr;
            lp -= pi * logp;
            p.dlogit[b * NA + a] = (pr * pi_sum - pi) * invB;
        }
        p.lossb[2 * b] = (v - t) * (v - t); p.lossb[2 * b + 1] = lp;
        red[0] = p.du[b];
    }
    __syncthreads();
    p.dh1[(size_t)b * HID + tid] = h > 0.f ? red[0] * p.w2[tid] : 0.f;
}

// gradient w.r.t. the head activations (through the Linear layers and Dropout), scattered into the gradient planes.
// A block takes 8 boards x 336 consecutive activations (grid: board groups x 8 activation groups): every weight it loads is used
// for 8 boards (one block per board re-read the 1.4 MB of fc_value 128 times: 138 us).
constexpr int HB_BOARDS = 8, HB_K = 336;
__global__ void __launch_bounds__(HB_K) head_bwd_x_kernel(const HeadP p) {
    pdl_sync();
    __shared__ float dh[HB_BOARDS][HID], dl[HB_BOARDS][8];
    const int b0 = blockIdx.x * HB_BOARDS, tid = threadIdx.x;
    for (int i = tid; i < HB_BOARDS * HID; i += HB_K) {
        const int b = b0 + i / HID;
        dh[i / HID][i % HID] = b < p.B ? p.dh1[(size_t)b * HID + i % HID] : 0.f;
    }
    if (tid < HB_BOARDS * 8) { const int b = b0 + tid / 8, a = tid % 8; dl[tid / 8][a] = (b < p.B && a < NA) ? p.dlogit[b * NA + a] : 0.f; }
    __syncthreads();
    const int k = blockIdx.y * HB_K + tid;           // 0 .. 2 * FLAT - 1
    const int head = k / FLAT, kk = k - head * FLAT, c = kk / CELLS + 32 * head, cell = kk % CELLS;
    float acc[HB_BOARDS];
#pragma unroll
    for (int i = 0; i < HB_BOARDS; ++i) acc[i] = 0.f;
    if (head == 0) {
        for (int a = 0; a < NA; ++a) {
            const float w = p.wp[a * FLAT + kk];
#pragma unroll
            for (int i = 0; i < HB_BOARDS; ++i) acc[i] = fmaf(dl[i][a], w, acc[i]);
        }
    } else {
#pragma unroll 4
        for (int j = 0; j < HID; ++j) {
            const float w = p.w1[(size_t)j * FLAT + kk];
#pragma unroll
            for (int i = 0; i < HB_BOARDS; ++i) acc[i] = fmaf(dh[i][j], w, acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < HB_BOARDS; ++i) {
        const int b = b0 + i;
        if (b >= p.B) break;
        const size_t o = ((size_t)b * 2 + head) * FLAT + kk;
        const float gx = p.mask[o] ? 2.f * acc[i] : 0.f;
        const int g = GUARD + b * BOARD_ROWS + (cell / 6) * 7 + cell % 6;
        p.gah[((size_t)(c >> 2) * p.Rg + g) * 4 + (c & 3)] = gx;
    }
}

// weight gradients of the Linear layers + the loss scalars.  One thread per output element, loop over the batch.
struct FcGradP {
    const float* xd; const float* dh1; const float* dlogit; const float* du; const float* h1; const float* lossb;
    float* g_w1; float* g_b1; float* g_wp; float* g_bp; float* g_w2; float* g_b2; float* loss_out;
    int B;
};
__global__ void __launch_bounds__(256) fc_wgrad_kernel(const FcGradP p) {
    pdl_sync();
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int n_w1 = HID * FLAT, n_wp = NA * FLAT;
    if (i < n_w1) {
        const int j = i / FLAT, k = i - j * FLAT;
        float acc = 0.f;
        for (int b = 0; b < p.B; ++b) acc = fmaf(p.dh1[(size_t)b * HID + j], p.xd[((size_t)b * 2 + 1) * FLAT + k], acc);
        p.g_w1[i] = acc;
    } else if (i < n_w1 + n_wp) {
        const int ii = i - n_w1, a = ii / FLAT, k = ii - a * FLAT;
        float acc = 0.f;
        for (int b = 0; b < p.B; ++b) acc = fmaf(p.dlogit[b * NA + a], p.xd[((size_t)b * 2) * FLAT + k], acc);
        p.g_wp[ii] = acc;
    } else if (i < n_w1 + n_wp + HID) {
        const int j = i - n_w1 - n_wp;
        float a1 = 0.f, a2 = 0.f;
        for (int b = 0; b < p.B; ++b) { a1 += p.dh1[(size_t)b * HID + j]; a2 = fmaf(p.du[b], p.h1[(size_t)b * HID + j], a2); }
        p.g_b1[j] = a1; p.g_w2[j] = a2;
    } else if (i < n_w1 + n_wp + HID + NA) {
        const int a = i - n_w1 - n_wp - HID;
        float acc = 0.f;
        for (int b = 0; b < p.B; ++b) acc += p.dlogit[b * NA + a];
        p.g_bp[a] = acc;
    } else if (i == n_w1 + n_wp + HID + NA) {
        float db2 = 0.f, lv = 0.f, lp = 0.f;
        for (int b = 0; b < p.B; ++b) { db2 += p.du[b]; lv += p.lossb[2 * b]; lp += p.lossb[2 * b + 1]; }
        p.g_b2[0] = db2;
        lv /= (float)p.B; lp /= (float)p.B;
        p.loss_out[0] = lv + lp; p.loss_out[1] = lv; p.loss_out[2] = lp;
    }
}

// ------------------------------------------------------------------------------------------------ SGD + weight packing
// torch.optim.SGD(momentum, weight_decay), dampening 0, no Nesterov: g += wd * p; buf = mu * buf + g; p -= lr * buf
__global__ void sgd_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, long long n, float lr, float mu, float wd) {
    pdl_sync();
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gg = fmaf(wd, p[i], g[i]);
    const float mm = fmaf(mu, m[i], gg);
    m[i] = mm;
    p[i] = fmaf(-lr, mm, p[i]);
}

// conv weights W[co][ci][tap] (two tensors split over co for the fused head conv) -> the two packed forms conv_tf32_kernel reads
// (a CTA computes 64 output channels, so the slices of each half are contiguous):
//   forward  wf[co / 64][tap][ks][kc][co % 64][i] = W[co][ci = 8 ks + 4 kc + i][tap]             (K = CINP input channels, zero beyond CIN)
//   backward wb[ci / 64][tap][ks][kc][ci % 64][i] = W[co = 8 ks + 4 kc + i][ci][taps - 1 - tap]   (K = COUT, N = CINP; absent for the stem)
__global__ void pack_conv_kernel(const float* __restrict__ w0, const float* __restrict__ w1, int split, int COUT, int CIN, int CINP, int taps, float* wf, float* wb) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= taps * CINP * COUT) return;
    const int q = i & 3, n64 = (i >> 2) & 63, kc = (i >> 8) & 1;
    {
        const int KS = CINP / 8, r = i >> 9, ks = r % KS, tap = (r / KS) % taps, half = r / (KS * taps);
        const int ci = 8 * ks + 4 * kc + q, co = half * 64 + n64;
        float v = 0.f;
        if (ci < CIN) v = co < split ? w0[((size_t)co * CIN + ci) * taps + tap] : w1[((size_t)(co - split) * CIN + ci) * taps + tap];
        wf[i] = v;
    }
    if (wb) {
        const int KS = COUT / 8, r = i >> 9, ks = r % KS, tap = (r / KS) % taps, half = r / (KS * taps);
        const int co = 8 * ks + 4 * kc + q, ci = half * 64 + n64, t = taps - 1 - tap;
        wb[i] = co < split ? w0[((size_t)co * CIN + ci) * taps + t] : w1[((size_t)(co - split) * CIN + ci) * taps + t];
    }
}
__global__ void transpose_w1_kernel(const float* __restrict__ w1, float* w1t) {
    pdl_sync();   // [256][1344] -> [1344][256]
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < HID * FLAT) { const int j = i / FLAT, k = i - j * FLAT; w1t[(size_t)k * HID + j] = w1[i]; }
}

}  // namespace train
}  // namespace spx

// ================================================================================================ host side
using namespace spx::train;

// every kernel of the trainer goes through here: programmatic stream serialization (see pdl_sync) + the launch counter
template <typename... KArgs, typename... Args>
static void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
    spx::count_launch();
}


struct ConvLayer {
    size_t w, b, gamma, beta;     // offsets into the flat parameter vector
    size_t rstat;                 // offset into the running-statistics vector (mean[N], var[N])
};

struct spx_trainer {
    int blocks, L, B, Bpad, R, Rg, tiles, n_real, S;
    size_t n_params, n_running;
    std::vector<ConvLayer> conv;                 // stem + 2 * blocks trunk layers
    size_t pol_w, pol_b, pol_g, pol_be, lp_w, lp_b, val_w, val_b, val_g, val_be, fc_w, fc_b, lo_w, lo_b, pol_rs, val_rs;
    float *params, *grads, *mom, *running;
    float *x0, *y, *a, *yh, *ah, *g0, *g1, *skip, *dy, *dyh, *gah;     // planes (y / a: L tensors back to back)
    float *mean, *invstd;                        // [(L + 1)][128]
    float *stat, *part, *partial;
    float *wf, *wb;                              // packed conv weights, L layers of 9 * 128 * 128 each (stem: 9 * 8 * 128 at the front of its slot)
    float *whf, *whb;                            // head conv packs
    float *w1t, *xd, *probs, *value, *h1, *dlogit, *du, *dh1, *lossb, *loss;
    unsigned char* mask;
    __nv_bfloat16 *a16, *dy16;                   // bf16 planes: activations of all L layers / one conv-output gradient
    size_t plane128, plane64;                    // floats per 128- / 64-channel plane tensor
    bool packed;
};

static size_t conv_w_elems(int l) { return l == 0 ? (size_t)CH * 3 * 9 : (size_t)CH * CH * 9; }

extern "C" {

/* Parameter order = ResidualTower.named_parameters() (general/modules.py:42-75): conv1.{weight,bias}, bn1.{weight,bias},
 * residual_blocks.i.{conv1,bn1,conv2,bn2}.{weight,bias}, conv_policy, policy_bn, linear_policy, conv_value, value_bn, fc_value,
 * linear_output.  Running statistics: per BatchNorm in module order, mean[C] then var[C]. */
int spx_train_create(int32_t num_blocks, int32_t batch, spx_trainer** out) {
    if (!out || num_blocks < 1 || num_blocks > 64 || batch < 2 || batch > 4096) return spx::set_err(SPX_E_ARG, "spx_train_create: bad argument%s", "");
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return spx::set_err(SPX_E_CUDA, "spx_train_create: no CUDA device%s", "");
    spx_trainer* t = new (std::nothrow) spx_trainer();
    if (!t) return spx::set_err(SPX_E_ARG, "spx_train_create: out of host memory%s", "");
    t->blocks = num_blocks; t->L = 2 * num_blocks + 1; t->B = batch;
    t->Bpad = (batch + 15) / 16 * 16;                       // 16 boards = 896 rows = 7 tiles
    t->R = t->Bpad * BOARD_ROWS; t->Rg = t->R + 2 * GUARD; t->tiles = t->R / TILE; t->n_real = batch * CELLS;
    t->S = t->tiles < 28 ? t->tiles : 28;
    size_t off = 0, rs = 0;
    for (int l = 0; l < t->L; ++l) {
        ConvLayer c;
        c.w = off; off += conv_w_elems(l);
        c.b = off; off += CH; c.gamma = off; off += CH; c.beta = off; off += CH;
        c.rstat = rs; rs += 2 * CH;
        t->conv.push_back(c);
    }
    t->pol_w = off; off += 32 * CH; t->pol_b = off; off += 32; t->pol_g = off; off += 32; t->pol_be = off; off += 32;
    t->lp_w = off; off += NA * FLAT; t->lp_b = off; off += NA;
    t->val_w = off; off += 32 * CH; t->val_b = off; off += 32; t->val_g = off; off += 32; t->val_be = off; off += 32;
    t->fc_w = off; off += (size_t)HID * FLAT; t->fc_b = off; off += HID; t->lo_w = off; off += HID; t->lo_b = off; off += 1;
    t->pol_rs = rs; rs += 64; t->val_rs = rs; rs += 64;
    t->n_params = off; t->n_running = rs;
    t->plane128 = (size_t)(CH / 4) * t->Rg * 4; t->plane64 = (size_t)(HEAD / 4) * t->Rg * 4;
    struct { float** p; size_t n; } allocs[] = {
        {&t->params, off}, {&t->grads, off}, {&t->mom, off}, {&t->running, rs},
        {&t->x0, (size_t)2 * t->Rg * 4}, {&t->y, t->plane128 * t->L}, {&t->a, t->plane128 * t->L}, {&t->yh, t->plane64}, {&t->ah, t->plane64},
        {&t->g0, t->plane128}, {&t->g1, t->plane128}, {&t->skip, t->plane128}, {&t->dy, t->plane128}, {&t->dyh, t->plane64}, {&t->gah, t->plane64},
        {&t->mean, (size_t)(t->L + 1) * CH}, {&t->invstd, (size_t)(t->L + 1) * CH},
        {&t->stat, (size_t)t->tiles * 2 * CH}, {&t->part, (size_t)t->tiles * 2 * CH}, {&t->partial, (size_t)t->S * 9 * CH * CH},
        {&t->wf, (size_t)t->L * 9 * CH * CH}, {&t->wb, (size_t)t->L * 9 * CH * CH}, {&t->whf, (size_t)CH * HEAD}, {&t->whb, (size_t)CH * HEAD},
        {&t->w1t, (size_t)HID * FLAT}, {&t->xd, (size_t)batch * 2 * FLAT}, {&t->probs, (size_t)batch * NA}, {&t->value, (size_t)batch},
        {&t->h1, (size_t)batch * HID}, {&t->dlogit, (size_t)batch * NA}, {&t->du, (size_t)batch}, {&t->dh1, (size_t)batch * HID},
        {&t->lossb, (size_t)batch * 2}, {&t->loss, 4}};
    for (auto& al : allocs) {
        if (cudaMalloc((void**)al.p, al.n * sizeof(float)) != cudaSuccess || cudaMemset(*al.p, 0, al.n * sizeof(float)) != cudaSuccess)
            return spx::set_err(SPX_E_CUDA, "spx_train_create: cudaMalloc failed%s", "");
    }
    if (cudaMalloc((void**)&t->mask, (size_t)batch * 2 * FLAT) != cudaSuccess) return spx::set_err(SPX_E_CUDA, "spx_train_create: cudaMalloc failed%s", "");
    const size_t n16 = (size_t)(CH / 8) * t->Rg * 8;
    if (cudaMalloc((void**)&t->a16, n16 * t->L * 2) != cudaSuccess || cudaMalloc((void**)&t->dy16, n16 * 2) != cudaSuccess ||
        cudaMemset(t->a16, 0, n16 * t->L * 2) != cudaSuccess || cudaMemset(t->dy16, 0, n16 * 2) != cudaSuccess)
        return spx::set_err(SPX_E_CUDA, "spx_train_create: cudaMalloc failed%s", "");
    SPX_CUDA_R(cudaFuncSetAttribute(conv_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, CONV_SMEM));
    SPX_CUDA_R(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    t->packed = false;
    *out = t;
    return 0;
}

int spx_train_destroy(spx_trainer* t) {
    if (!t) return 0;
    float* ptrs[] = {t->params, t->grads, t->mom, t->running, t->x0, t->y, t->a, t->yh, t->ah, t->g0, t->g1, t->skip, t->dy, t->dyh, t->gah, t->mean,
                     t->invstd, t->stat, t->part, t->partial, t->wf, t->wb, t->whf, t->whb, t->w1t, t->xd, t->probs, t->value, t->h1, t->dlogit, t->du,
                     t->dh1, t->lossb, t->loss};
    for (float* p : ptrs) if (p) cudaFree(p);
    if (t->mask) cudaFree(t->mask);
    if (t->a16) cudaFree(t->a16);
    if (t->dy16) cudaFree(t->dy16);
    delete t;
    return 0;
}

int64_t spx_train_param_count(spx_trainer* t) { return t ? (int64_t)t->n_params : 0; }
int64_t spx_train_running_count(spx_trainer* t) { return t ? (int64_t)t->n_running : 0; }

static int repack(spx_trainer* t, cudaStream_t st) {
    for (int l = 0; l < t->L; ++l) {
        const int cin = l == 0 ? 3 : CH, cinp = l == 0 ? 8 : CH;
        const int n = 9 * cinp * CH;
        launch_k(pack_conv_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, t->params + t->conv[l].w, nullptr, CH, CH, cin, cinp, 9, t->wf + (size_t)l * 9 * CH * CH,
                                                             l == 0 ? nullptr : t->wb + (size_t)l * 9 * CH * CH);
    }
    launch_k(pack_conv_kernel, dim3((unsigned)((CH * HEAD + 255) / 256)), dim3(256), 0, st, t->params + t->pol_w, t->params + t->val_w, 32, HEAD, CH, CH, 1, t->whf, t->whb);
    launch_k(transpose_w1_kernel, dim3((unsigned)((HID * FLAT + 255) / 256)), dim3(256), 0, st, t->params + t->fc_w, t->w1t);
    SPX_CUDA_R(cudaGetLastError());
    t->packed = true;
    return 0;
}

/* copy parameters / running statistics in (device pointers; either may be NULL) and rebuild the packed weights;
 * reset_momentum != 0 clears the momentum buffers (a fresh torch.optim.SGD) */
int spx_train_set_state(spx_trainer* t, const float* dev_params, const float* dev_running, int32_t reset_momentum, void* stream) {
    if (!t) return spx::set_err(SPX_E_ARG, "spx_train_set_state: null trainer%s", "");
    cudaStream_t st = (cudaStream_t)stream;
    if (dev_params) SPX_CUDA_R(cudaMemcpyAsync(t->params, dev_params, t->n_params * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (dev_running) SPX_CUDA_R(cudaMemcpyAsync(t->running, dev_running, t->n_running * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (reset_momentum) SPX_CUDA_R(cudaMemsetAsync(t->mom, 0, t->n_params * sizeof(float), st));
    return repack(t, st);
}

/* what: 0 parameters, 1 running statistics, 2 gradients of the last step, 3 momentum buffers (device destinations) */
int spx_train_get_state(spx_trainer* t, int32_t what, float* dev_out, void* stream) {
    if (!t || !dev_out) return spx::set_err(SPX_E_ARG, "spx_train_get_state: null argument%s", "");
    const float* src = what == 0 ? t->params : what == 1 ? t->running : what == 2 ? t->grads : what == 3 ? t->mom : nullptr;
    if (!src) return spx::set_err(SPX_E_ARG, "spx_train_get_state: unknown selector%s", "");
    const size_t n = what == 1 ? t->n_running : t->n_params;
    SPX_CUDA_R(cudaMemcpyAsync(dev_out, src, n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

/* network outputs of the last step's forward pass: probs [B][7], value [B] (device destinations, either may be NULL) */
int spx_train_outputs(spx_trainer* t, float* dev_probs, float* dev_value, void* stream) {
    if (!t) return spx::set_err(SPX_E_ARG, "spx_train_outputs: null trainer%s", "");
    if (dev_probs) SPX_CUDA_R(cudaMemcpyAsync(dev_probs, t->probs, (size_t)t->B * NA * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    if (dev_value) SPX_CUDA_R(cudaMemcpyAsync(dev_value, t->value, (size_t)t->B * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return 0;
}

/* test hook: device pointer and size (floats) of an internal plane tensor [C/4][rows][4] (rows = 16 + 56 * padded batch).
 * which: 0 x0, 1 y[layer], 2 a[layer], 3 yh, 4 ah, 5 g0, 6 g1, 7 dy, 8 dyh, 9 gah, 10 skip */
int spx_train_debug_planes(spx_trainer* t, int32_t which, int32_t layer, float** dev_ptr, int64_t* n_floats, int32_t* rows) {
    if (!t || !dev_ptr || !n_floats || !rows || layer < 0 || layer >= t->L) return spx::set_err(SPX_E_ARG, "spx_train_debug_planes: bad argument%s", "");
    float* p = nullptr; size_t n = 0;
    switch (which) {
        case 0: p = t->x0; n = (size_t)2 * t->Rg * 4; break;
        case 1: p = t->y + (size_t)layer * t->plane128; n = t->plane128; break;
        case 2: p = t->a + (size_t)layer * t->plane128; n = t->plane128; break;
        case 3: p = t->yh; n = t->plane64; break;
        case 4: p = t->ah; n = t->plane64; break;
        case 5: p = t->g0; n = t->plane128; break;
        case 6: p = t->g1; n = t->plane128; break;
        case 7: p = t->dy; n = t->plane128; break;
        case 8: p = t->dyh; n = t->plane64; break;
        case 9: p = t->gah; n = t->plane64; break;
        case 10: p = t->skip; n = t->plane128; break;
        default: return spx::set_err(SPX_E_ARG, "spx_train_debug_planes: unknown tensor%s", "");
    }
    *dev_ptr = p; *n_floats = (int64_t)n; *rows = t->Rg;
    return 0;
}

static void wg_default_strides(WgP& wp) { wp.a_lbo = 128; wp.a_sbo = PIECE; wp.b_lbo = 128; wp.b_sbo = TILE * 16; }

/* test hook: the backward-weights kernel on caller-provided bf16 plane tensors [C/8][rows][8] (x: 128 channels, dy: N channels, `rows` rows each incl.
 * the 16 guard rows, rows - 16 a multiple of 128); partial_out f32 [S][taps][128][N].  lbo/sbo < 0: the product's strides. */
int spx_train_debug_wgrad(const void* x, const void* dy, int32_t N, int32_t taps, int32_t rows, int32_t S, float* partial_out,
                          int32_t a_lbo, int32_t a_sbo, int32_t b_lbo, int32_t b_sbo, void* stream) {
    if (!x || !dy || !partial_out || (N != 64 && N != 128) || (taps != 1 && taps != 9) || (rows - 16) % TILE || S < 1 || S > (rows - 16) / TILE)
        return spx::set_err(SPX_E_ARG, "spx_train_debug_wgrad: bad argument%s", "");
    SPX_CUDA_R(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    WgP wp; wp.x = (const __nv_bfloat16*)x; wp.dy = (const __nv_bfloat16*)dy; wp.N = N; wp.taps = taps; wp.partial = partial_out; wp.Rg = rows; wp.nchunks = (rows - 16) / TILE;
    wg_default_strides(wp);
    if (a_lbo >= 0) { wp.a_lbo = a_lbo; wp.a_sbo = a_sbo; wp.b_lbo = b_lbo; wp.b_sbo = b_sbo; }
    launch_k(wgrad_kernel, dim3(S, taps == 9 ? 3 : 1), dim3(256), WG_SMEM, (cudaStream_t)stream, wp);
    SPX_CUDA_R(cudaGetLastError());
    return 0;
}

static long long* g_conv_trace = nullptr;
/* test hook: device buffer of 8 int64 that receives clock64() at 8 points of CTA 0 of every conv_tf32_kernel launch (NULL: off) */
int spx_train_debug_trace(long long* dev_trace) { g_conv_trace = dev_trace; return 0; }

static Seg2 seg(float* p) { Seg2 s; s.p0 = p; s.p1 = p; s.split = 1 << 30; return s; }
static Seg2 seg2(float* p0, float* p1) { Seg2 s; s.p0 = p0; s.p1 = p1; s.split = 32; return s; }

static void launch_conv(const ConvP& p0, int tiles, cudaStream_t st) {
    ConvP p = p0;
    p.trace = g_conv_trace;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(tiles, p.N / NC); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = CONV_SMEM; cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    attr[0].id = cudaLaunchAttributeClusterDimension;      // 1, or 7 consecutive row tiles (the tile count is a multiple of 7) sharing the weight stream
    static int cl = 0;
    if (!cl) { const char* e = getenv("SPX_CONV_CLUSTER"); cl = e ? atoi(e) : CONV_CLUSTER; if (cl != 1 && cl != 7) cl = CONV_CLUSTER; }
    attr[0].val.clusterDim.x = cl; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 2;
    cudaLaunchKernelEx(&cfg, conv_tf32_kernel, p);
    spx::count_launch();
}

/* One update_from_memory step (mcts.py:254-270) on a batch already on the device:
 *   planes f32[B][3][7][6] (preprocess of the sampled states, modules.py:115-125; spx_replay_sample writes them),
 *   tree_probs f32[B][7], target f32[B] (actual_val, + q when q_average: mcts.py:243-244),
 *   dropout_mask u8[B][2][1344] or NULL (NULL: keep-masks drawn from (seed, step); [.][0] policy head, [.][1] value head),
 *   apply_update 0: forward + backward only (gradients readable with spx_train_get_state(2)).
 * loss_out (device, 3 floats, may be NULL): total, value term, policy term.  Asynchronous on `stream`. */
int spx_train_step(spx_trainer* t, const float* planes, const float* tree_probs, const float* target, const uint8_t* dropout_mask,
                   uint64_t seed, uint64_t step, float lr, float momentum, float weight_decay, int32_t apply_update, float* loss_out, void* stream) {
    if (!t || !planes || !tree_probs || !target) return spx::set_err(SPX_E_ARG, "spx_train_step: null argument%s", "");
    if (!t->packed) return spx::set_err(SPX_E_STATE, "spx_train_step: call spx_train_set_state first%s", "");
    cudaStream_t st = (cudaStream_t)stream;
    const int L = t->L, B = t->B, Rg = t->Rg, tiles = t->tiles;
    float* P = t->params; float* G = t->grads;
    launch_k(input_kernel, dim3((unsigned)((Rg + 255) / 256)), dim3(256), 0, st, planes, t->x0, B, Rg);
    // ---------------- forward
    for (int l = 0; l < L; ++l) {
        const ConvLayer& c = t->conv[l];
        ConvP cp; memset(&cp, 0, sizeof(cp));
        cp.in = l == 0 ? t->x0 : t->a + (size_t)(l - 1) * t->plane128; cp.in_chunks = l == 0 ? 2 : CH / 4;
        cp.w = t->wf + (size_t)l * 9 * CH * CH; cp.taps = 9; cp.ks_per_tap = l == 0 ? 1 : CH / 8; cp.N = CH;
        cp.out = t->y + (size_t)l * t->plane128; cp.bias = seg(P + c.b); cp.has_bias = 1; cp.add = nullptr; cp.stat = t->stat; cp.B = B; cp.Rg = Rg;
        launch_conv(cp, tiles, st);
        BnP bp; memset(&bp, 0, sizeof(bp));
        bp.y = cp.out; bp.a = t->a + (size_t)l * t->plane128;
        bp.res = (l >= 2 && (l & 1) == 0) ? t->a + (size_t)(l - 2) * t->plane128 : nullptr;     // out += identity (modules.py:37)
        bp.stat = t->stat; bp.gamma = seg(P + c.gamma); bp.beta = seg(P + c.beta);
        bp.rmean = seg(t->running + c.rstat); bp.rvar = seg(t->running + c.rstat + CH);
        bp.mean = t->mean + (size_t)l * CH; bp.invstd = t->invstd + (size_t)l * CH;
        bp.a16 = t->a16 + (size_t)l * (CH / 8) * Rg * 8;
        bp.N = CH; bp.B = B; bp.Rg = Rg; bp.tiles = tiles; bp.n_real = t->n_real;
        launch_k(bn_fwd_kernel, dim3(tiles, bp.N / BN_CG), dim3(256), 0, st, bp);
    }
    const float* a_last = t->a + (size_t)(L - 1) * t->plane128;
    {   // fused policy | value 1x1 head conv + its two BatchNorms + ReLU (modules.py:97,102)
        ConvP cp; memset(&cp, 0, sizeof(cp));
        cp.in = a_last; cp.in_chunks = CH / 4; cp.w = t->whf; cp.taps = 1; cp.ks_per_tap = CH / 8; cp.N = HEAD; cp.out = t->yh;
        cp.bias = seg2(P + t->pol_b, P + t->val_b); cp.has_bias = 1; cp.stat = t->stat; cp.B = B; cp.Rg = Rg;
        launch_conv(cp, tiles, st);
        BnP bp; memset(&bp, 0, sizeof(bp));
        bp.y = t->yh; bp.a = t->ah; bp.stat = t->stat; bp.gamma = seg2(P + t->pol_g, P + t->val_g); bp.beta = seg2(P + t->pol_be, P + t->val_be);
        bp.rmean = seg2(t->running + t->pol_rs, t->running + t->val_rs); bp.rvar = seg2(t->running + t->pol_rs + 32, t->running + t->val_rs + 32);
        bp.mean = t->mean + (size_t)L * CH; bp.invstd = t->invstd + (size_t)L * CH;
        bp.N = HEAD; bp.B = B; bp.Rg = Rg; bp.tiles = tiles; bp.n_real = t->n_real;
        launch_k(bn_fwd_kernel, dim3(tiles, bp.N / BN_CG), dim3(256), 0, st, bp);
    }
    HeadP hp; memset(&hp, 0, sizeof(hp));
    hp.ah = t->ah; hp.mask_in = dropout_mask; hp.mask = t->mask; hp.xd = t->xd; hp.wp = P + t->lp_w; hp.bp = P + t->lp_b; hp.w1t = t->w1t; hp.b1 = P + t->fc_b;
    hp.w1 = P + t->fc_w; hp.w2 = P + t->lo_w; hp.b2 = P + t->lo_b; hp.tree_probs = tree_probs; hp.target = target; hp.probs = t->probs; hp.value = t->value;
    hp.h1 = t->h1; hp.dlogit = t->dlogit; hp.du = t->du; hp.dh1 = t->dh1; hp.lossb = t->lossb; hp.gah = t->gah; hp.seed = seed; hp.step = step; hp.B = B; hp.Rg = Rg;
    launch_k(head_fwd_kernel, dim3((unsigned)(B)), dim3(256), 0, st, hp);
    // ---------------- backward: heads
    launch_k(head_bwd_x_kernel, dim3((unsigned)((B + HB_BOARDS - 1) / HB_BOARDS), 2 * FLAT / HB_K), dim3(HB_K), 0, st, hp);
    FcGradP fg; fg.xd = t->xd; fg.dh1 = t->dh1; fg.dlogit = t->dlogit; fg.du = t->du; fg.h1 = t->h1; fg.lossb = t->lossb;
    fg.g_w1 = G + t->fc_w; fg.g_b1 = G + t->fc_b; fg.g_wp = G + t->lp_w; fg.g_bp = G + t->lp_b; fg.g_w2 = G + t->lo_w; fg.g_b2 = G + t->lo_b;
    fg.loss_out = t->loss; fg.B = B;
    launch_k(fc_wgrad_kernel, dim3((unsigned)((HID * FLAT + NA * FLAT + HID + NA + 1 + 255) / 256)), dim3(256), 0, st, fg);
    {
        BnP bp; memset(&bp, 0, sizeof(bp));
        bp.y = t->yh; bp.a = t->ah; bp.g = t->gah; bp.dy = t->dyh; bp.dy16 = t->dy16; bp.skip = nullptr; bp.part = t->part;
        bp.gamma = seg2(P + t->pol_g, P + t->val_g); bp.mean = t->mean + (size_t)L * CH; bp.invstd = t->invstd + (size_t)L * CH;
        bp.dgamma = seg2(G + t->pol_g, G + t->val_g); bp.dbeta = seg2(G + t->pol_be, G + t->val_be); bp.dbias = seg2(G + t->pol_b, G + t->val_b);
        bp.N = HEAD; bp.B = B; bp.Rg = Rg; bp.tiles = tiles; bp.n_real = t->n_real;
        launch_k(bn_bwd_reduce_kernel, dim3(tiles, bp.N / BN_CG), dim3(256), 0, st, bp);
        launch_k(bn_bwd_apply_kernel, dim3(tiles, bp.N / BN_CG), dim3(256), 0, st, bp);
        WgP wp; wp.x = t->a16 + (size_t)(L - 1) * (CH / 8) * Rg * 8; wp.dy = t->dy16; wp.N = HEAD; wp.taps = 1; wp.partial = t->partial; wp.Rg = Rg; wp.nchunks = tiles; wg_default_strides(wp);
        launch_k(wgrad_kernel, dim3(t->S, 1), dim3(256), WG_SMEM, st, wp);
        launch_k(wgrad_reduce_kernel, dim3((unsigned)((CH * HEAD + 255) / 256)), dim3(256), 0, st, t->partial, t->S, 1, CH, CH, HEAD, G + t->pol_w, G + t->val_w, 32);
        ConvP cp; memset(&cp, 0, sizeof(cp));    // gradient w.r.t. the trunk output: dyh (64 channels) x W^T
        cp.in = t->dyh; cp.in_chunks = HEAD / 4; cp.w = t->whb; cp.taps = 1; cp.ks_per_tap = HEAD / 8; cp.N = CH; cp.out = t->g0; cp.has_bias = 0; cp.bias = seg(nullptr);
        cp.B = B; cp.Rg = Rg;
        launch_conv(cp, tiles, st);
    }
    // ---------------- backward: trunk
    float* gcur = t->g0; float* gnext = t->g1;
    for (int l = L - 1; l >= 0; --l) {
        const ConvLayer& c = t->conv[l];
        const bool second = l >= 2 && (l & 1) == 0, first = (l & 1) == 1;
        BnP bp; memset(&bp, 0, sizeof(bp));
        bp.y = t->y + (size_t)l * t->plane128; bp.a = t->a + (size_t)l * t->plane128; bp.g = gcur; bp.dy = t->dy; bp.dy16 = t->dy16; bp.skip = second ? t->skip : nullptr; bp.part = t->part;
        bp.gamma = seg(P + c.gamma); bp.mean = t->mean + (size_t)l * CH; bp.invstd = t->invstd + (size_t)l * CH;
        bp.dgamma = seg(G + c.gamma); bp.dbeta = seg(G + c.beta); bp.dbias = seg(G + c.b);
        bp.N = CH; bp.B = B; bp.Rg = Rg; bp.tiles = tiles; bp.n_real = t->n_real;
        launch_k(bn_bwd_reduce_kernel, dim3(tiles, bp.N / BN_CG), dim3(256), 0, st, bp);
        launch_k(bn_bwd_apply_kernel, dim3(tiles, bp.N / BN_CG), dim3(256), 0, st, bp);
        if (l == 0) {
            const int splits = t->S, rows_per = (t->R + splits - 1) / splits;
            launch_k(stem_wgrad_kernel, dim3(splits, 9), dim3(CH), 0, st, t->x0, t->dy, t->partial, Rg, t->R, rows_per);
            launch_k(wgrad_reduce_kernel, dim3((unsigned)((9 * 3 * CH + 255) / 256)), dim3(256), 0, st, t->partial, splits, 9, 3, 3, CH, G + c.w, G + c.w, 1 << 30);
            break;
        }
        WgP wp; wp.x = t->a16 + (size_t)(l - 1) * (CH / 8) * Rg * 8; wp.dy = t->dy16; wp.N = CH; wp.taps = 9; wp.partial = t->partial; wp.Rg = Rg; wp.nchunks = tiles; wg_default_strides(wp);
        launch_k(wgrad_kernel, dim3(t->S, 3), dim3(256), WG_SMEM, st, wp);
        launch_k(wgrad_reduce_kernel, dim3((unsigned)((9 * CH * CH + 255) / 256)), dim3(256), 0, st, t->partial, t->S, 9, CH, CH, CH, G + c.w, G + c.w, 1 << 30);
        ConvP cp; memset(&cp, 0, sizeof(cp));
        cp.in = t->dy; cp.in_chunks = CH / 4; cp.w = t->wb + (size_t)l * 9 * CH * CH; cp.taps = 9; cp.ks_per_tap = CH / 8; cp.N = CH; cp.out = gnext;
        cp.has_bias = 0; cp.bias = seg(nullptr); cp.add = first ? t->skip : nullptr; cp.B = B; cp.Rg = Rg;
        launch_conv(cp, tiles, st);
        float* tmp = gcur; gcur = gnext; gnext = tmp;
    }
    if (loss_out) SPX_CUDA_R(cudaMemcpyAsync(loss_out, t->loss, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (apply_update) {
        launch_k(sgd_kernel, dim3((unsigned)((unsigned)((t->n_params + 255) / 256))), dim3(256), 0, st, P, G, t->mom, (long long)t->n_params, lr, momentum, weight_decay);
        if (repack(t, st)) return SPX_E_CUDA;
    }
    SPX_CUDA_R(cudaGetLastError());
    return 0;
}

}  // extern "C"

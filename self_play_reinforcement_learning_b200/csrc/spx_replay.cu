// spx_replay.cu -- device-resident replay memory: the reference's Memory (rl_utils/memory.py:8-30) and the batch assembly of
// MCTreeSearch.loss (games/algos/mcts.py:234-243) without the records ever leaving HBM.
//
//   ring    spx_record[P]   physical ring; logical record k (0 = oldest) lives at (head + k) % P, size <= max_size <= P
//   append  deque(maxlen).append x n   (one coalesced copy kernel; eviction is pure host bookkeeping)
//   sample  np.random.choice(size, batch, replace=False) restated as a partial Fisher-Yates over the logical index space with
//           draws from the counter stream (oracle/replay.py is the CPU twin; integer only => bit-identical), then one gather
//           kernel that expands the bitboards into Move.state boards / preprocess planes and copies the targets.
// HBM-bound byte shuffling: 80 B read per sampled record, (8 + 12) * W*H + 4*(A + 2) B written.
#include <cuda_runtime.h>
#include <stdint.h>

#include <new>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "spx_common.cuh"

namespace spx {
int set_err(int code, const char* fmt, const char* detail);
void count_launch();
namespace replay {

constexpr int PURPOSE_SAMPLE = 4;
constexpr int MAX_BATCH = 4096;
constexpr uint32_t EMPTY = 0xFFFFFFFFu;

__global__ void append_kernel(spx_record* __restrict__ ring, long long P, long long start, const spx_record* __restrict__ src, long long n) {
    // 80-byte records moved as 5 x 16 B per record so that consecutive threads touch consecutive 16 B words
    const long long words = n * 5;
    const uint4* s = reinterpret_cast<const uint4*>(src);
    uint4* d = reinterpret_cast<uint4*>(ring);
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < words; w += (long long)gridDim.x * blockDim.x) {
        const long long rec = w / 5, part = w - rec * 5;
        d[((start + rec) % P) * 5 + part] = s[w];
    }
}

// One CTA: (1) every thread draws its share of the batch's random words, (2) thread 0 runs the partial Fisher-Yates with
// the displaced entries of the virtual array [0, size) in a shared-memory hash table, (3) indices go to global memory.
__global__ void __launch_bounds__(256) index_kernel(long long size, int batch, u64 seed, u64 step, int log2cap, long long* __restrict__ idx_out) {
    extern __shared__ __align__(16) unsigned char smem[];
    u64* draws = reinterpret_cast<u64*>(smem);
    uint32_t* keys = reinterpret_cast<uint32_t*>(draws + batch);
    const uint32_t cap = 1u << log2cap;
    uint32_t* vals = keys + cap;
    const u64 prefix = rng_prefix(seed, step, 0, PURPOSE_SAMPLE, 0);
    for (int i = threadIdx.x; i < batch; i += blockDim.x) draws[i] = sm64(sm64(prefix ^ (u64)(uint32_t)i));   // == spec.rng_u64(..., sim=i, depth=0, idx=0)
    for (uint32_t i = threadIdx.x; i < cap; i += blockDim.x) keys[i] = EMPTY;
    __syncthreads();
    if (threadIdx.x != 0) return;
    auto slot_of = [&](uint32_t k) {
        uint32_t h = (k * 2654435761u) >> (32 - log2cap);
        while (keys[h] != EMPTY && keys[h] != k) h = (h + 1) & (cap - 1);
        return h;
    };
    for (int i = 0; i < batch; ++i) {
        const uint32_t j = (uint32_t)(i + (long long)(draws[i] % (u64)(size - i)));
        const uint32_t si = slot_of((uint32_t)i);
        const uint32_t vi = keys[si] == EMPTY ? (uint32_t)i : vals[si];
        const uint32_t sj = slot_of(j);
        const uint32_t vj = keys[sj] == EMPTY ? j : vals[sj];
        idx_out[i] = (long long)vj;
        keys[sj] = j; vals[sj] = vi;     // a[j] = a[i]; a[i] is never read again
    }
}

template <int GAME>
__global__ void __launch_bounds__(64) gather_kernel(const spx_record* __restrict__ ring, long long P, long long head, const long long* __restrict__ idx,
                                                    long long* __restrict__ boards, float* __restrict__ planes, float* __restrict__ tree_probs,
                                                    float* __restrict__ actual_val, float* __restrict__ q) {
    using R = Rules<GAME>;
    const long long b = blockIdx.x;
    const spx_record* rec = ring + (head + idx[b]) % P;
    const int c = threadIdx.x;
    if (c < R::CELLS) {
        const int bit = (c / R::H) * R::STRIDE + (c % R::H);
        const int own = (int)((rec->own >> bit) & 1ULL), opp = (int)((rec->opp >> bit) & 1ULL);
        if (boards) boards[b * R::CELLS + c] = (long long)(own - opp);
        if (planes) {   // preprocess (general/modules.py:115-125): [== 0, == +1, == -1]
            float* p = planes + b * 3 * R::CELLS + c;
            p[0] = (float)(1 - own - opp); p[R::CELLS] = (float)own; p[2 * R::CELLS] = (float)opp;
        }
    }
    if (tree_probs && c < R::A) tree_probs[b * R::A + c] = rec->tree_probs[c];
    if (c == 0) {
        if (actual_val) actual_val[b] = rec->actual_val;
        if (q) q[b] = rec->q;
    }
}


// ------------------------------------------------------------------------------------------------ Deduplicator (memory.py:56-94)
// The reference keeps a dict keyed by the state bytes: per key the running SUMS of the value fields and a count, in
// first-insertion order; create_memory emits sum / count per key.  Here the dict is a device table of spx_record (sums in the
// value fields, count in pad1) in first-seen order; folding n new records into it is: stable sort of (table ++ new) by
// (own, opp) [two LSD radix passes], segment heads, first-seen position = exclusive scan of "is head" in ORIGINAL order, and one
// thread per segment that adds the members sequentially in insertion order (the same f32 additions, in the same order, as
// count[value] += experience.value).
struct DedupSrc {
    const spx_record* table; long long n_table;
    const spx_record* pending;
    __device__ __forceinline__ const spx_record& at(long long i) const { return i < n_table ? table[i] : pending[i - n_table]; }
};

__global__ void dedup_key_kernel(DedupSrc src, long long n, int which, const unsigned* __restrict__ order, u64* __restrict__ keys, unsigned* __restrict__ idx) {
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        const long long i = order ? (long long)order[j] : j;
        keys[j] = which == 0 ? src.at(i).opp : src.at(i).own;
        if (!order) idx[j] = (unsigned)j;
    }
}

__device__ __forceinline__ bool dedup_is_head(const DedupSrc& src, const unsigned* order, long long j) {
    if (j == 0) return true;
    const spx_record& a = src.at(order[j]);
    const spx_record& b = src.at(order[j - 1]);
    return a.own != b.own || a.opp != b.opp;
}

__global__ void dedup_head_kernel(DedupSrc src, long long n, const unsigned* __restrict__ order, int* __restrict__ first_flag) {
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x)
        first_flag[order[j]] = dedup_is_head(src, order, j) ? 1 : 0;   // stable sort: the head is the first-seen member
}

__global__ void dedup_fold_kernel(DedupSrc src, long long n, const unsigned* __restrict__ order, const int* __restrict__ pos, spx_record* __restrict__ out) {
    for (long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
        if (!dedup_is_head(src, order, j)) continue;
        const long long first = order[j];
        spx_record acc = src.at(first);
        if (first >= src.n_table) acc.pad1 = 1;                       // a raw record opens a new entry with count 1
        for (long long k = j + 1; k < n && !dedup_is_head(src, order, k); ++k) {
            const spx_record& m = src.at(order[k]);                   // members after the head are raw records, in insertion order
#pragma unroll
            for (int a = 0; a < SPX_MAX_ACTIONS; ++a) acc.tree_probs[a] = __fadd_rn(acc.tree_probs[a], m.tree_probs[a]);
            acc.actual_val = __fadd_rn(acc.actual_val, m.actual_val);
            acc.q = __fadd_rn(acc.q, m.q);
            acc.pad1 += 1;
        }
        out[pos[first]] = acc;
    }
}

// create_memory: ring[i] = table[first + i] with every value field divided by the count (tensor / int, true division)
__global__ void dedup_emit_kernel(const spx_record* __restrict__ table, long long first, long long keep, spx_record* __restrict__ ring) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < keep; i += (long long)gridDim.x * blockDim.x) {
        spx_record r = table[first + i];
        const float c = (float)r.pad1;
#pragma unroll
        for (int a = 0; a < SPX_MAX_ACTIONS; ++a) r.tree_probs[a] = __fdiv_rn(r.tree_probs[a], c);
        r.actual_val = __fdiv_rn(r.actual_val, c);
        r.q = __fdiv_rn(r.q, c);
        ring[i] = r;
    }
}

__global__ void ring_linearise_kernel(const spx_record* __restrict__ ring, long long P, long long head, long long n, spx_record* __restrict__ out) {
    const long long words = n * 5;
    const uint4* s = reinterpret_cast<const uint4*>(ring);
    uint4* d = reinterpret_cast<uint4*>(out);
    for (long long w = blockIdx.x * (long long)blockDim.x + threadIdx.x; w < words; w += (long long)gridDim.x * blockDim.x) {
        const long long rec = w / 5, part = w - rec * 5;
        d[w] = s[((head + rec) % P) * 5 + part];
    }
}

static inline unsigned blocks_for(long long n) {
    long long b = (n + 255) / 256;
    return (unsigned)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

}  // namespace replay
}  // namespace spx

struct spx_replay {
    spx_record* ring;
    long long* idx_scratch;
    int64_t P, M, head, size;
    // Deduplicator state (memory.py:56-62): `table` = counter dict in first-seen order, `pending` = temp_queue
    bool dedup_active = false;
    spx_record* table = nullptr;   int64_t n_table = 0;
    spx_record* pending = nullptr; int64_t n_pending = 0, cap_pending = 0;
};

#define RP_CUDA(expr)                                                                                    \
    do {                                                                                                 \
        cudaError_t e_ = (expr);                                                                         \
        if (e_ != cudaSuccess) return spx::set_err(SPX_E_CUDA, "spx_replay: %s", cudaGetErrorString(e_)); \
    } while (0)

static int pending_reserve(spx_replay* r, int64_t need, cudaStream_t st) {
    if (need <= r->cap_pending) return 0;
    int64_t cap = r->cap_pending ? r->cap_pending : 1024;
    while (cap < need) cap *= 2;
    spx_record* p = nullptr;
    RP_CUDA(cudaMalloc((void**)&p, sizeof(spx_record) * (size_t)cap));
    if (r->n_pending) RP_CUDA(cudaMemcpyAsync(p, r->pending, sizeof(spx_record) * (size_t)r->n_pending, cudaMemcpyDeviceToDevice, st));
    RP_CUDA(cudaStreamSynchronize(st));
    cudaFree(r->pending);
    r->pending = p; r->cap_pending = cap;
    return 0;
}

extern "C" {

int spx_replay_create(int64_t max_size, int64_t physical_capacity, spx_replay** out) {
    if (!out || max_size < 1 || physical_capacity < max_size || physical_capacity >= (1LL << 31))
        return spx::set_err(SPX_E_ARG, "spx_replay_create: need 1 <= max_size <= physical_capacity < 2^31%s", "");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return spx::set_err(SPX_E_CUDA, "spx_replay_create: no CUDA device (there is no CPU fallback)%s", "");
    spx_replay* r = new (std::nothrow) spx_replay();
    if (!r) return spx::set_err(SPX_E_ARG, "spx_replay_create: out of host memory%s", "");
    r->P = physical_capacity; r->M = max_size; r->head = 0; r->size = 0;
    if (cudaMalloc((void**)&r->ring, sizeof(spx_record) * (size_t)r->P) != cudaSuccess ||
        cudaMalloc((void**)&r->idx_scratch, sizeof(long long) * spx::replay::MAX_BATCH) != cudaSuccess) {
        cudaFree(r->ring); delete r;
        return spx::set_err(SPX_E_CUDA, "spx_replay_create: cudaMalloc failed%s", "");
    }
    cudaFuncSetAttribute(spx::replay::index_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * spx::replay::MAX_BATCH + 8 * 4 * spx::replay::MAX_BATCH);
    *out = r;
    return 0;
}

int spx_replay_destroy(spx_replay* r) {
    if (!r) return 0;
    cudaFree(r->ring); cudaFree(r->idx_scratch); cudaFree(r->table); cudaFree(r->pending);
    delete r;
    return 0;
}

int64_t spx_replay_size(spx_replay* r) { return r ? r->size : -1; }
int64_t spx_replay_max_size(spx_replay* r) { return r ? r->M : -1; }

int spx_replay_change_size(spx_replay* r, int64_t max_size) {
    if (!r || max_size < 1 || max_size > r->P) return spx::set_err(SPX_E_ARG, "spx_replay_change_size: max_size must be in [1, physical_capacity]%s", "");
    if (r->size > max_size) { r->head = (r->head + (r->size - max_size)) % r->P; r->size = max_size; }   // deque(buffer, maxlen): the newest survive
    r->M = max_size;
    return 0;
}

int spx_replay_reset(spx_replay* r) {
    if (!r) return spx::set_err(SPX_E_ARG, "spx_replay_reset: null%s", "");
    r->head = 0; r->size = 0;
    return 0;
}

int spx_replay_append(spx_replay* r, const spx_record* dev_records, int64_t n, void* stream) {
    if (!r || n < 0 || (n > 0 && !dev_records)) return spx::set_err(SPX_E_ARG, "spx_replay_append: bad argument%s", "");
    if (n == 0) return 0;
    if (r->dedup_active) {   // Memory.add: self.deduplicator.add_temp(experience) (memory.py:19-20) -- every record, evicted or not
        if (int rc = pending_reserve(r, r->n_pending + n, (cudaStream_t)stream)) return rc;
        RP_CUDA(cudaMemcpyAsync(r->pending + r->n_pending, dev_records, sizeof(spx_record) * (size_t)n, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        r->n_pending += n;
    }
    const int64_t keep = n < r->M ? n : r->M, skip = n - keep;   // a deque(maxlen=M) fed n > M items keeps the last M
    const int64_t start = (r->head + r->size) % r->P;
    long long blocks = (keep * 5 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    spx::replay::append_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(r->ring, r->P, start, dev_records + skip, keep);
    spx::count_launch();
    RP_CUDA(cudaGetLastError());
    const int64_t total = r->size + keep, evict = total > r->M ? total - r->M : 0;
    r->head = (r->head + evict) % r->P;
    r->size = total - evict;
    return 0;
}

int spx_replay_read(spx_replay* r, int64_t first, int64_t n, spx_record* host_out, void* stream) {
    if (!r || first < 0 || n < 0 || first + n > r->size || (n > 0 && !host_out)) return spx::set_err(SPX_E_ARG, "spx_replay_read: range outside [0, size)%s", "");
    if (n == 0) return 0;
    const int64_t p0 = (r->head + first) % r->P, n0 = (p0 + n <= r->P) ? n : r->P - p0;
    RP_CUDA(cudaMemcpyAsync(host_out, r->ring + p0, sizeof(spx_record) * n0, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    if (n0 < n) RP_CUDA(cudaMemcpyAsync(host_out + n0, r->ring, sizeof(spx_record) * (n - n0), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    RP_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return 0;
}

int spx_replay_sample(spx_replay* r, int32_t game, int64_t batch, uint64_t seed, uint64_t step, int64_t* idx, int64_t* boards, float* planes,
                      float* tree_probs, float* actual_val, float* q, void* stream) {
    if (!r || (game != SPX_GAME_CONNECT4 && game != SPX_GAME_TICTACTOE)) return spx::set_err(SPX_E_ARG, "spx_replay_sample: bad argument%s", "");
    if (batch < 1 || batch > r->size || batch > spx::replay::MAX_BATCH)
        return spx::set_err(SPX_E_STATE, "spx_replay_sample: need 1 <= batch <= min(len(memory), 4096) (mcts.py:255-261 skips the update instead)%s", "");
    cudaStream_t st = (cudaStream_t)stream;
    long long* ix = idx ? (long long*)idx : r->idx_scratch;
    int log2cap = 4;
    while ((1LL << log2cap) < 2 * batch) ++log2cap;
    const size_t smem = 8 * (size_t)batch + 8 * ((size_t)1 << log2cap);
    spx::replay::index_kernel<<<1, 256, smem, st>>>(r->size, (int)batch, seed, step, log2cap, ix);
    spx::count_launch();
    if (game == SPX_GAME_CONNECT4)
        spx::replay::gather_kernel<SPX_GAME_CONNECT4><<<(unsigned)batch, 64, 0, st>>>(r->ring, r->P, r->head, ix, (long long*)boards, planes, tree_probs, actual_val, q);
    else
        spx::replay::gather_kernel<SPX_GAME_TICTACTOE><<<(unsigned)batch, 64, 0, st>>>(r->ring, r->P, r->head, ix, (long long*)boards, planes, tree_probs, actual_val, q);
    spx::count_launch();
    RP_CUDA(cudaGetLastError());
    return 0;
}

int64_t spx_replay_unique(spx_replay* r) { return r ? r->n_table : -1; }

int spx_replay_deduplicate(spx_replay* r, int64_t maxlen, void* stream) {
    using namespace spx::replay;
    if (!r) return spx::set_err(SPX_E_ARG, "spx_replay_deduplicate: null%s", "");
    cudaStream_t st = (cudaStream_t)stream;
    if (!r->dedup_active) {   // Deduplicator(buffer=self._buffer): the first call folds the whole current buffer (memory.py:48-49,62)
        if (int rc = pending_reserve(r, r->size, st)) return rc;
        if (r->size) {
            ring_linearise_kernel<<<blocks_for(r->size * 5), 256, 0, st>>>(r->ring, r->P, r->head, r->size, r->pending);
            spx::count_launch();
        }
        r->n_pending = r->size;
        r->dedup_active = true;
    }
    const int64_t n = r->n_table + r->n_pending;
    const int64_t max_eff = (maxlen <= 0 || maxlen > r->P) ? r->P : maxlen;   // deque(maxlen=None) is bounded by the physical ring here
    if (n == 0) { r->head = 0; r->size = 0; r->M = max_eff; return 0; }
    if (n >= (1LL << 31)) return spx::set_err(SPX_E_OVERFLOW, "spx_replay_deduplicate: more than 2^31 entries%s", "");
    if (r->n_pending > 0) {
        u64 *keys_a = nullptr, *keys_b = nullptr; unsigned *idx_a = nullptr, *idx_b = nullptr; int *flag = nullptr, *pos = nullptr;
        spx_record* fresh = nullptr; void* tmp = nullptr;
        size_t tmp_sort = 0, tmp_scan = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tmp_sort, keys_a, keys_b, idx_a, idx_b, (int)n, 0, 64, st);
        cub::DeviceScan::ExclusiveSum(nullptr, tmp_scan, flag, pos, (int)n, st);
        const size_t tmp_bytes = tmp_sort > tmp_scan ? tmp_sort : tmp_scan;
        auto release = [&]() { cudaFree(keys_a); cudaFree(keys_b); cudaFree(idx_a); cudaFree(idx_b); cudaFree(flag); cudaFree(pos); cudaFree(tmp); };
        if (cudaMalloc((void**)&keys_a, 8 * n) != cudaSuccess || cudaMalloc((void**)&keys_b, 8 * n) != cudaSuccess ||
            cudaMalloc((void**)&idx_a, 4 * n) != cudaSuccess || cudaMalloc((void**)&idx_b, 4 * n) != cudaSuccess ||
            cudaMalloc((void**)&flag, 4 * n) != cudaSuccess || cudaMalloc((void**)&pos, 4 * n) != cudaSuccess ||
            cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 16) != cudaSuccess || cudaMalloc((void**)&fresh, sizeof(spx_record) * (size_t)n) != cudaSuccess) {
            release(); cudaFree(fresh);
            return spx::set_err(SPX_E_CUDA, "spx_replay_deduplicate: cudaMalloc failed%s", "");
        }
        const DedupSrc src{r->table, r->n_table, r->pending};
        const unsigned g = blocks_for(n);
        size_t tb = tmp_bytes;
        dedup_key_kernel<<<g, 256, 0, st>>>(src, n, 0, nullptr, keys_a, idx_a);                       // LSD pass 1: by opp
        cub::DeviceRadixSort::SortPairs(tmp, tb, keys_a, keys_b, idx_a, idx_b, (int)n, 0, 64, st);
        dedup_key_kernel<<<g, 256, 0, st>>>(src, n, 1, idx_b, keys_a, nullptr);                       // LSD pass 2: by own (stable)
        tb = tmp_bytes;
        cub::DeviceRadixSort::SortPairs(tmp, tb, keys_a, keys_b, idx_b, idx_a, (int)n, 0, 64, st);    // idx_a = final order
        dedup_head_kernel<<<g, 256, 0, st>>>(src, n, idx_a, flag);
        tb = tmp_bytes;
        cub::DeviceScan::ExclusiveSum(tmp, tb, flag, pos, (int)n, st);
        dedup_fold_kernel<<<g, 256, 0, st>>>(src, n, idx_a, pos, fresh);
        for (int k = 0; k < 4; ++k) spx::count_launch();
        int last_pos = 0, last_flag = 0;
        cudaError_t e1 = cudaMemcpyAsync(&last_pos, pos + (n - 1), 4, cudaMemcpyDeviceToHost, st);
        cudaError_t e2 = cudaMemcpyAsync(&last_flag, flag + (n - 1), 4, cudaMemcpyDeviceToHost, st);
        cudaError_t e3 = cudaStreamSynchronize(st);
        cudaError_t e4 = cudaGetLastError();
        release();
        if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess || e4 != cudaSuccess) {
            cudaFree(fresh);
            return spx::set_err(SPX_E_CUDA, "spx_replay_deduplicate: %s", cudaGetErrorString(e3 != cudaSuccess ? e3 : (e4 != cudaSuccess ? e4 : (e1 != cudaSuccess ? e1 : e2))));
        }
        cudaFree(r->table);
        r->table = fresh; r->n_table = (int64_t)last_pos + last_flag; r->n_pending = 0;
    }
    // create_memory(max_size): a deque(maxlen) fed every entry in first-seen order keeps the last maxlen (memory.py:86-94)
    const int64_t keep = r->n_table < max_eff ? r->n_table : max_eff;
    dedup_emit_kernel<<<blocks_for(keep), 256, 0, st>>>(r->table, r->n_table - keep, keep, r->ring);
    spx::count_launch();
    RP_CUDA(cudaGetLastError());
    r->head = 0; r->size = keep; r->M = max_eff;
    return 0;
}

}  // extern "C"

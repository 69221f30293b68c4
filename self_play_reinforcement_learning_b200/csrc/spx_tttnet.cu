// spx_tttnet.cu -- the repo's TicTacToe network (games/tictactoe/modules.py:14-81, ConvNetTicTacToe) in fp32.
//
// 3 x (conv3x3 + BN + leaky_relu 0.01): 3 -> 128 -> 128 -> 64 on the 3x3 board, 1x1 head convs (2 policy / 1 value channel)
// + BN + leaky, Linear(18 -> A) + softmax, Linear(9 -> 256) + leaky -> Linear(256 -> 1) + tanh.  4 MFLOP per leaf: far too
// small for the tensor cores (a 3x3 board is 9 GEMM rows), so this is a plain fp32 CUDA-core kernel, one CTA per board with
// all activations in shared memory, BN folded into the convolutions on the host.  fp32 throughout => it is compared with the
// fp32 torch forward at 1e-5 (tests/test_tttnet_gpu.py).  Used by BASELINE.json configs[0] ("repo's tictactoe net").
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <new>

#include "spx_common.cuh"

namespace spx {
int set_err(int code, const char* fmt, const char* detail);
void count_launch();
namespace ttt {

constexpr int C1 = 128, C2 = 128, C3 = 64, CELLS = 9, A = 9, HID = 256;
// blob layout (floats); conv weights are stored [ic][tap][oc] so that thread == output channel reads coalesced
constexpr int OFF_W1 = 0, OFF_B1 = OFF_W1 + 3 * 9 * C1;
constexpr int OFF_W2 = OFF_B1 + C1, OFF_B2 = OFF_W2 + C1 * 9 * C2;
constexpr int OFF_W3 = OFF_B2 + C2, OFF_B3 = OFF_W3 + C2 * 9 * C3;
constexpr int OFF_WP = OFF_B3 + C3, OFF_BP = OFF_WP + C3 * 2;   // policy 1x1 conv [ic][2]
constexpr int OFF_WV = OFF_BP + 2, OFF_BV = OFF_WV + C3;        // value 1x1 conv [ic]
constexpr int OFF_LP = OFF_BV + 1, OFF_LPB = OFF_LP + A * 18;   // linear_policy [A][18]
constexpr int OFF_F1 = OFF_LPB + A, OFF_F1B = OFF_F1 + HID * 9; // fc_value [256][9]
constexpr int OFF_F2 = OFF_F1B + HID, OFF_F2B = OFF_F2 + HID;   // linear_output [256]
constexpr int BLOB_FLOATS = OFF_F2B + 1;

__device__ __forceinline__ float leaky(float x) { return x > 0.f ? x : 0.01f * x; }

// out[oc][pos] = leaky(b[oc] + sum_ic sum_tap w[ic][tap][oc] * in[ic][pos + tap]) for one thread's output channel
template <int CIN, int COUT>
__device__ __forceinline__ void conv3x3(const float* __restrict__ w, const float* __restrict__ b, const float (*in)[CELLS], float (*out)[CELLS], int oc) {
    if (oc >= COUT) return;
    float acc[CELLS];
#pragma unroll
    for (int p = 0; p < CELLS; ++p) acc[p] = b[oc];
    for (int ic = 0; ic < CIN; ++ic) {
        float x[CELLS];
#pragma unroll
        for (int p = 0; p < CELLS; ++p) x[p] = in[ic][p];
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
            const float wv = __ldg(w + ((size_t)ic * 9 + tap) * COUT + oc);
            const int dx = tap / 3 - 1, dy = tap % 3 - 1;   // tensor is [C, x, y]; kernel index (kh, kw) <-> (dx, dy)
#pragma unroll
            for (int p = 0; p < CELLS; ++p) {
                const int px = p / 3 + dx, py = p % 3 + dy;
                if (px >= 0 && px < 3 && py >= 0 && py < 3) acc[p] = fmaf(wv, x[px * 3 + py], acc[p]);
            }
        }
    }
#pragma unroll
    for (int p = 0; p < CELLS; ++p) out[oc][p] = leaky(acc[p]);
}

__global__ void __launch_bounds__(128) tttnet_kernel(const float* __restrict__ blob, const unsigned long long* __restrict__ own,
                                                     const unsigned long long* __restrict__ opp, const unsigned char* __restrict__ needs,
                                                     long long n, float* __restrict__ policy, float* __restrict__ value) {
    const long long b = blockIdx.x;
    if (b >= n || (needs && !needs[b])) return;
    __shared__ float x0[3][CELLS], x1[C1][CELLS], x2[C2][CELLS], x3[C3][CELLS], ph[18], vh[9], hid[HID], logit[A];
    const int t = threadIdx.x;
    if (t < CELLS) {   // preprocess (modules.py:56-66): planes (empty, own == +1, enemy == -1); bit index = x*3 + y
        const unsigned o = (unsigned)((own[b] >> t) & 1ULL), e = (unsigned)((opp[b] >> t) & 1ULL);
        x0[0][t] = (float)(1u - o - e); x0[1][t] = (float)o; x0[2][t] = (float)e;
    }
    __syncthreads();
    conv3x3<3, C1>(blob + OFF_W1, blob + OFF_B1, x0, x1, t);
    __syncthreads();
    conv3x3<C1, C2>(blob + OFF_W2, blob + OFF_B2, x1, x2, t);
    __syncthreads();
    conv3x3<C2, C3>(blob + OFF_W3, blob + OFF_B3, x2, x3, t);
    __syncthreads();
    if (t < 27) {   // 1x1 head convs + BN + leaky: channels 0,1 = policy, 2 = value; flatten index c*9 + x*3 + y
        const int c = t / 9, p = t % 9;
        float acc = c < 2 ? blob[OFF_BP + c] : blob[OFF_BV];
        for (int ic = 0; ic < C3; ++ic) acc = fmaf(c < 2 ? blob[OFF_WP + ic * 2 + c] : blob[OFF_WV + ic], x3[ic][p], acc);
        if (c < 2) ph[c * 9 + p] = leaky(acc); else vh[p] = leaky(acc);
    }
    __syncthreads();
    if (t < A) {
        float acc = blob[OFF_LPB + t];
        for (int k = 0; k < 18; ++k) acc = fmaf(blob[OFF_LP + t * 18 + k], ph[k], acc);
        logit[t] = acc;
    }
    for (int j = t; j < HID; j += 128) {
        float acc = blob[OFF_F1B + j];
        for (int k = 0; k < 9; ++k) acc = fmaf(blob[OFF_F1 + j * 9 + k], vh[k], acc);
        hid[j] = leaky(acc);
    }
    __syncthreads();
    if (t == 0) {
        float m = logit[0];
        for (int a = 1; a < A; ++a) m = fmaxf(m, logit[a]);
        float z = 0.f, e[A];
        for (int a = 0; a < A; ++a) { e[a] = expf(logit[a] - m); z += e[a]; }
        for (int a = 0; a < A; ++a) policy[b * A + a] = e[a] / z;
    }
    if (t < 32) {
        float acc = 0.f;
        for (int j = t; j < HID; j += 32) acc = fmaf(blob[OFF_F2 + j], hid[j], acc);
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (t == 0) value[b] = tanhf(acc + blob[OFF_F2B]);
    }
}

}  // namespace ttt
}  // namespace spx

struct spx_tttnet {
    float* blob;
};

extern "C" {

int64_t spx_tttnet_blob_floats(void) { return spx::ttt::BLOB_FLOATS; }

int spx_tttnet_create(spx_tttnet** out) {
    if (!out) return spx::set_err(SPX_E_ARG, "spx_tttnet_create: null out%s", "");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return spx::set_err(SPX_E_CUDA, "spx_tttnet_create: no CUDA device (there is no CPU fallback)%s", "");
    spx_tttnet* t = new (std::nothrow) spx_tttnet();
    if (!t) return spx::set_err(SPX_E_ARG, "spx_tttnet_create: out of host memory%s", "");
    if (cudaMalloc((void**)&t->blob, sizeof(float) * spx::ttt::BLOB_FLOATS) != cudaSuccess) { delete t; return spx::set_err(SPX_E_CUDA, "spx_tttnet_create: cudaMalloc failed%s", ""); }
    *out = t;
    return 0;
}

int spx_tttnet_destroy(spx_tttnet* t) {
    if (!t) return 0;
    cudaFree(t->blob);
    delete t;
    return 0;
}

int spx_tttnet_load(spx_tttnet* t, const float* dev_blob, int64_t n_floats, void* stream) {
    if (!t || !dev_blob || n_floats != spx::ttt::BLOB_FLOATS) return spx::set_err(SPX_E_ARG, "spx_tttnet_load: bad argument%s", "");
    cudaError_t e = cudaMemcpyAsync(t->blob, dev_blob, sizeof(float) * n_floats, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (e != cudaSuccess) return spx::set_err(SPX_E_CUDA, "spx_tttnet_load: %s", cudaGetErrorString(e));
    return 0;
}

int spx_tttnet_forward(spx_tttnet* t, const uint64_t* own, const uint64_t* opp, const uint8_t* needs_eval, int64_t n, float* policy,
                       float* value, void* stream) {
    if (!t || !own || !opp || !policy || !value) return spx::set_err(SPX_E_ARG, "spx_tttnet_forward: null argument%s", "");
    if (n <= 0) return 0;
    spx::ttt::tttnet_kernel<<<(unsigned)n, 128, 0, (cudaStream_t)stream>>>(t->blob, (const unsigned long long*)own, (const unsigned long long*)opp,
                                                                           needs_eval, n, policy, value);
    spx::count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return spx::set_err(SPX_E_CUDA, "spx_tttnet_forward: %s", cudaGetErrorString(e));
    return 0;
}

}  // extern "C"

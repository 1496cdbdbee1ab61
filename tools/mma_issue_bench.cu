// mma_issue_bench.cu -- how many cycles does one tcgen05.mma kind::tf32 cost when a single thread
// issues them back to back?  (Design input for csrc/tc_igemm.cuh: the GEMM kernels ran at
// ~110 + 0.3 N cycles per MMA where the math needs N / 2.)
//
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I fincflow_b200/csrc -o tools/bin/mma_issue_bench tools/mma_issue_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "../fincflow_b200/csrc/tc_common.cuh"

using namespace finc;
using namespace finc::tc;

// mode: 0 = TS (A from TMEM), constant operands;  1 = TS, operands advance like the GEMM k-loop;
//       2 = SS (A from shared memory), constant operands;  3 = TS, 4 rotating accumulators
template <int N, int MODE>
__global__ void __launch_bounds__(128, 1) bench_kernel(long long* out, int iters) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar, bar2, bar3, bar4, bar5, bar6;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 0.f;
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        mbar_init(&bar2, 1);
        mbar_init(&bar3, 1);
        mbar_init(&bar4, 1); mbar_init(&bar5, 1); mbar_init(&bar6, 1);
        mbar_arrive(&bar4); mbar_arrive(&bar5); mbar_arrive(&bar6);
        fence_mbar_init();
    }
    if (threadIdx.x < 32) tmem_alloc(&slot, 512);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = umma_idesc_tf32(128, N);
        const uint64_t db = umma_desc_k_sw128(smem), da = umma_desc_k_sw128(smem + 48 * 1024);
        const uint32_t ta = tmem + 256;
        long long t0 = clock64();
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int u = 0; u < 12; ++u) {
                if (MODE == 0) umma_tf32_ts(tmem, ta, db, idesc, 1);
                if (MODE == 1) umma_tf32_ts(tmem, ta + (u & 3) * 8 + (u >= 4 && u < 8 ? 32 : 0), db + (uint64_t)((u & 3) * 2), idesc, (i | u) != 0);
                if (MODE == 2) umma_tf32(tmem, da, db, idesc, 1);
                if (MODE == 3) umma_tf32_ts(tmem + (u & 3) * (N <= 64 ? N : 0), ta, db, idesc, 1);
                if (MODE == 4 || MODE == 5) umma_tf32_ts(tmem, ta + (u & 3) * 8 + (u >= 4 && u < 8 ? 32 : 0), db + (uint64_t)((u & 3) * 2), idesc, (i | u) != 0);
                if (u == 6 && MODE == 12) {   // look-ahead in the middle of the k-block: two waits polled together + fence
                    while (!(mbar_try_wait(&bar4, 0) & mbar_try_wait(&bar5, 0))) {}
                    tc_fence_after();
                }
                if (u == 6 && MODE == 15) { mbar_wait_long(&bar4, 0); }
                if (u == 3 && MODE == 16) { mbar_wait_long(&bar4, 0); }
                if (u == 9 && MODE == 16) { mbar_wait_long(&bar5, 0); tc_fence_after(); }
                if (MODE >= 6) {   // like the GEMM: ring of 4 stages (48 KB apart in smem, 64 columns apart in TMEM), 2 partials, acc = 0 at group start
                    const int st = i & 3;
                    const uint64_t dbs = umma_desc_k_sw128(smem + (st & 1) * 48 * 1024 + (u >= 4 && u < 8 ? 16 * 1024 : 0));
                    const uint32_t tas = tmem + 256 + st * 64;
                    umma_tf32_ts(tmem + ((i >> 1) & 1) * 128, tas + (u & 3) * 8 + (u < 4 ? 32 : 0), dbs + (uint64_t)((u & 3) * 2), idesc, ((i & 1) | u) != 0);
                }
            }
            if ((MODE == 4 || MODE == 5 || MODE >= 7) && MODE != 11) umma_commit(&bar2);
            if (MODE == 8 || MODE == 9) {   // a pause in the issuing thread after the commit
                const long long t = clock64();
                while (clock64() - t < (MODE == 8 ? 100 : 250)) {}
            }
            if (MODE == 10) {   // the GEMM's bookkeeping: three waits on completed barriers + fence
                mbar_wait_long(&bar4, 0); mbar_wait_long(&bar5, 0); mbar_wait_long(&bar6, 0);
                tc_fence_after();
            }
            if (MODE == 13) {
                while (!(mbar_try_wait(&bar4, 0) & mbar_try_wait(&bar5, 0))) {}
                tc_fence_after();
            }
            if (MODE == 14) tc_fence_after();
            if (MODE == 11) {   // pause BEFORE the commit (i.e. between MMAs only)
                const long long t = clock64();
                while (clock64() - t < 250) {}
            }
            if (MODE == 5 || (MODE == 7 && (i & 1))) umma_commit(&bar3);
        }
        long long t1 = clock64();
        umma_commit(&bar);
        mbar_wait_long(&bar, 0);
        long long t2 = clock64();
        if (blockIdx.x == 0) {
            out[0] = t1 - t0;   // issue time
            out[1] = t2 - t0;   // until the last MMA has completed
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int MODE>
static void run(const char* what, long long* d_out) {
    const int iters = 200;
    auto k = bench_kernel<N, MODE>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<<<148, 128, 100 * 1024>>>(d_out, iters);
    k<<<148, 128, 100 * 1024>>>(d_out, iters);
    cudaError_t err = cudaDeviceSynchronize();
    long long h[2] = {0, 0};
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    const double n = 12.0 * iters;
    printf("%-44s N=%3d: issue %7.1f cyc/MMA, complete %7.1f cyc/MMA (math floor %5.1f)  %s\n", what, N, h[0] / n, h[1] / n,
           N / 2.0, err == cudaSuccess ? "" : cudaGetErrorString(err));
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 16);
    run<16, 0>("TS, constant operands", d_out);
    run<48, 0>("TS, constant operands", d_out);
    run<128, 0>("TS, constant operands", d_out);
    run<256, 0>("TS, constant operands", d_out);
    run<16, 1>("TS, advancing operands (GEMM k-loop)", d_out);
    run<128, 1>("TS, advancing operands (GEMM k-loop)", d_out);
    run<16, 2>("SS, constant operands", d_out);
    run<128, 2>("SS, constant operands", d_out);
    run<256, 2>("SS, constant operands", d_out);
    run<16, 3>("TS, 4 rotating accumulators", d_out);
    run<48, 3>("TS, 4 rotating accumulators", d_out);
    run<128, 4>("TS, advancing, 1 commit per 12 MMAs", d_out);
    run<128, 5>("TS, advancing, 2 commits per 12 MMAs", d_out);
    run<128, 6>("TS, GEMM-like rings, no commits", d_out);
    run<128, 7>("TS, GEMM-like rings + commits", d_out);
    run<128, 8>("... + 100-cycle pause after the commit", d_out);
    run<128, 9>("... + 250-cycle pause after the commit", d_out);
    run<128, 10>("... + 3 completed mbarrier waits + fence", d_out);
    run<128, 11>("... 250-cycle pause, no commit", d_out);
    run<128, 12>("... 2 waits polled together + fence, mid-block", d_out);
    run<128, 13>("... 2 waits polled together + fence, at the end", d_out);
    run<128, 14>("... fence only", d_out);
    run<128, 15>("... 1 wait mid-block", d_out);
    run<128, 16>("... 1 wait after MMA 3, 1 wait + fence after MMA 9", d_out);
    return 0;
}

// finc_collective.cu -- gradient all-reduce over NVLink peer memory fused with the Adam update.
//
// Training is the only place the FInC path exchanges data between GPUs: the flat masked
// gradient bucket (436 KB for the CIFAR-shaped flow) must be summed over the ranks before the
// optimiser step.  Instead of NCCL all-reduce -> Adam (two launches, a ring/tree protocol sized
// for large messages), every rank runs ONE kernel that
//   1. announces "my bucket is complete" to all peers and waits for theirs (flags in the peers'
//      signal pads, st.release.sys / ld.acquire.sys through NVLink),
//   2. reads element e of every peer's bucket straight from peer memory (P2P loads over
//      NVLink/NVSwitch, fixed rank order => bit-identical sums and parameters on all ranks),
//   3. applies Adam to its local replica of the parameters,
//   4. announces "done reading" so that peers may overwrite their buckets.
// The buckets and signal pads are symmetric allocations (torch.distributed._symmetric_memory);
// the kernel only sees raw pointers.  One-shot (every rank reads all buckets) is the right
// algorithm for a message this small: (W-1) * 436 KB per rank is ~4 us of NVLink time.
#include "finc_common.cuh"

namespace finc {

namespace {

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

constexpr int kMaxWorld = 16;

struct CollArgs {
    const float* const* peer_grad;  // [world] device pointers to every rank's bucket (symmetric)
    unsigned* const* peer_signal;   // [world] device pointers to every rank's signal pad (>= 2*kMaxWorld u32)
    unsigned* local;                // local scratch: [0] epoch, [1] go flag, [2] done counter
    float* param;
    float* exp_avg;
    float* exp_avg_sq;
    float* step;
    float lr, b1, b2, eps, grad_scale;
    long n;
    int rank, world;
};

// spin with a bound: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void wait_sys(const unsigned* p, unsigned v) {
    for (unsigned long long spins = 0; ld_acquire_sys(p) != v; ++spins)
        if (spins > (1ull << 31)) __trap();
}

__global__ void __launch_bounds__(256) allreduce_adam_kernel(const CollArgs a) {
    __shared__ unsigned s_epoch;
    const unsigned nblk = gridDim.x;
    if (threadIdx.x == 0) s_epoch = ld_acquire_gpu(&a.local[0]) + 1;
    __syncthreads();
    const unsigned epoch = s_epoch;

    // ---- 1. arrival barrier across ranks (CTA 0), then release the other CTAs ----------------------
    if (blockIdx.x == 0) {
        if (threadIdx.x < a.world) {
            __threadfence_system();  // this rank's bucket (written by earlier kernels) -> visible to peers
            st_release_sys(a.peer_signal[threadIdx.x] + a.rank, epoch);
            wait_sys(a.peer_signal[a.rank] + threadIdx.x, epoch);
        }
        __syncthreads();
        if (threadIdx.x == 0) st_release_gpu(&a.local[1], epoch);
    } else {
        if (threadIdx.x == 0) {
            for (unsigned long long spins = 0; ld_acquire_gpu(&a.local[1]) != epoch; ++spins)
                if (spins > (1ull << 31)) __trap();
        }
        __syncthreads();
    }

    // ---- 2 + 3. sum the peers' buckets in rank order, Adam on the local replica ------------------
    const float t = *a.step + 1.f;
    const float c1 = 1.f - powf(a.b1, t), c2 = 1.f - powf(a.b2, t);
    const float* gp[kMaxWorld];
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r) gp[r] = r < a.world ? a.peer_grad[r] : nullptr;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < a.n; e += (long)nblk * blockDim.x) {
        float gs[kMaxWorld];
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r) gs[r] = r < a.world ? __ldcg(gp[r] + e) : 0.f;  // all loads in flight
        float g = 0.f;
#pragma unroll
        for (int r = 0; r < kMaxWorld; ++r) g += gs[r];
        g *= a.grad_scale;
        const float me = a.b1 * a.exp_avg[e] + (1.f - a.b1) * g;
        const float ve = a.b2 * a.exp_avg_sq[e] + (1.f - a.b2) * g * g;
        a.exp_avg[e] = me;
        a.exp_avg_sq[e] = ve;
        a.param[e] -= a.lr * (me / c1) / (sqrtf(ve / c2) + a.eps);
    }

    // ---- 4. departure: every CTA of this rank is done reading -> tell the peers ----------------------
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&a.local[2], 1u);
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            for (unsigned long long spins = 0; ld_acquire_gpu(&a.local[2]) != nblk; ++spins)
                if (spins > (1ull << 31)) __trap();
            a.local[2] = 0u;
            *a.step = t;
        }
        __syncthreads();
        if (threadIdx.x < a.world) {
            st_release_sys(a.peer_signal[threadIdx.x] + kMaxWorld + a.rank, epoch);
            wait_sys(a.peer_signal[a.rank] + kMaxWorld + threadIdx.x, epoch);
        }
        __syncthreads();
        if (threadIdx.x == 0) st_release_gpu(&a.local[0], epoch);
    }
}

}  // namespace

int launch_allreduce_adam(const void* peer_grad, const void* peer_signal, void* local, float* param, float* m, float* v,
                          float* step, float lr, float b1, float b2, float eps, float grad_scale, long n, int rank,
                          int world, cudaStream_t st) {
    if (world < 1 || world > kMaxWorld) return FINC_E_BADARG;
    CollArgs a{};
    a.peer_grad = static_cast<const float* const*>(peer_grad);
    a.peer_signal = static_cast<unsigned* const*>(const_cast<void*>(peer_signal));
    a.local = static_cast<unsigned*>(local);
    a.param = param; a.exp_avg = m; a.exp_avg_sq = v; a.step = step;
    a.lr = lr; a.b1 = b1; a.b2 = b2; a.eps = eps; a.grad_scale = grad_scale; a.n = n; a.rank = rank; a.world = world;
    long blocks = (n + 255) / 256;
    const int sms = sm_count_cached();
    if (blocks > sms) blocks = sms;  // all CTAs must be co-resident (they wait on CTA 0)
    if (blocks < 1) blocks = 1;
    allreduce_adam_kernel<<<(unsigned)blocks, 256, 0, st>>>(a);
    return (int)cudaGetLastError();
}

}  // namespace finc

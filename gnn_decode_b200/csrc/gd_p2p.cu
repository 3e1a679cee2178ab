// One-shot all-reduce of the flat gradient vector over NVLink / NVSwitch PEER MEMORY, fused into a single
// kernel (the training path's only exchange step is 10h+3 = 1 283 floats -- latency-bound, so a
// ring or tree is pointless: every rank publishes its vector in its own symmetric buffer, raises an epoch flag, waits
// for the peers' flags and sums all vectors in RANK ORDER straight from peer memory: the same fixed order on every
// rank -> bit-identical results everywhere).  The symmetric buffers and their peer mappings come from
// torch.distributed._symmetric_memory (plumbing); the data path is this kernel.
//   buffer of rank r (floats):  [ slot 0: n_pad ][ slot 1: n_pad ][ flag (uint32, as one float slot) ]
// Double-buffered by epoch parity: a rank overwrites slot p at step k+2 only after passing the flag wait of step
// k+1, by which time every peer has finished reading slot p of step k.
#include "gd_adam.cuh"

namespace gd {

constexpr int kMaxWorld = 16;
struct P2PParams {
    float* peer[kMaxWorld];       // peer[r] = rank r's symmetric buffer mapped into this process
    const float* src;
    float* dst;
    int* err;                     // set to 1 when a peer never showed up (bounded spin); dst and the weights are then left untouched
    int world, rank, n, n_pad;
    unsigned int epoch;
    float scale;
    // optional fused optimizer step on the reduced gradient (gd_p2p_allreduce_adam): every rank applies the same
    // bit-identical update to its replica of the weights, so the replicas never drift
    float *w, *m, *v;
    AdamCoef adam;
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void __launch_bounds__(256) p2p_allreduce_kernel(const P2PParams p) {
    __shared__ int timed_out;
    const int tid = threadIdx.x;
    if (tid == 0) timed_out = 0;
    float* mine = p.peer[p.rank] + (size_t)(p.epoch & 1u) * p.n_pad;
    for (int i = tid; i < p.n; i += blockDim.x) mine[i] = p.src[i];
    __threadfence_system();
    __syncthreads();
    if (tid == 0) st_release_sys(reinterpret_cast<unsigned int*>(p.peer[p.rank] + 2 * (size_t)p.n_pad), p.epoch);
    if (tid < p.world && tid != p.rank) {
        const unsigned int* flag = reinterpret_cast<const unsigned int*>(p.peer[tid] + 2 * (size_t)p.n_pad);
        long long spins = 0;
        // epochs only grow; (int) difference tolerates wrap-around
        while ((int)(ld_acquire_sys(flag) - p.epoch) < 0) {
            // 2^24 polls of >= 64 ns each plus the system-scope load: tens of seconds.  A peer is gone: do not hang the
            // GPU, and do NOT touch dst / the weights with a partial sum (the flag is CTA-uniform after the barrier)
            if (++spins > (1ll << 24)) { *p.err = 1; timed_out = 1; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (timed_out) return;
    for (int i = tid; i < p.n; i += blockDim.x) {
        float acc = 0.f;
        for (int r = 0; r < p.world; ++r)                             // fixed rank order on every rank
            acc += __ldcg(p.peer[r] + (size_t)(p.epoch & 1u) * p.n_pad + i);
        acc *= p.scale;
        if (p.dst) p.dst[i] = acc;
        if (p.w) {
            float wi = p.w[i], mi = p.m[i], vi = p.v[i];
            adam_update(p.adam, acc, wi, mi, vi);
            p.w[i] = wi; p.m[i] = mi; p.v[i] = vi;
        }
    }
}

}  // namespace gd

extern "C" int64_t gd_p2p_buffer_floats(int32_t n) {
    if (n <= 0) return -1;
    return 2 * (int64_t)((n + 31) / 32 * 32) + 32;
}

extern "C" int gd_p2p_allreduce(const uint64_t* peer_ptrs_host, int32_t world, int32_t rank, const float* src_dev,
                                float* dst_dev, int32_t n, uint32_t epoch, float scale, int32_t* err_dev, void* stream) {
    GD_CHECK_ARG(peer_ptrs_host && src_dev && dst_dev && err_dev, "gd_p2p_allreduce: NULL argument");
    GD_CHECK_ARG(world >= 1 && world <= gd::kMaxWorld && rank >= 0 && rank < world, "gd_p2p_allreduce: bad rank %d / world %d",
                 rank, world);
    GD_CHECK_ARG(n > 0 && epoch > 0, "gd_p2p_allreduce: n and epoch must be positive");
    gd::P2PParams p;
    for (int r = 0; r < world; ++r) p.peer[r] = reinterpret_cast<float*>((uintptr_t)peer_ptrs_host[r]);
    p.src = src_dev; p.dst = dst_dev; p.err = err_dev; p.world = world; p.rank = rank; p.n = n; p.n_pad = (n + 31) / 32 * 32;
    p.epoch = epoch; p.scale = scale; p.w = p.m = p.v = nullptr;
    gd::p2p_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p);
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

extern "C" int gd_p2p_allreduce_adam(const uint64_t* peer_ptrs_host, int32_t world, int32_t rank, const float* src_dev,
                                     float* dst_dev, int32_t n, uint32_t epoch, float scale, int32_t* err_dev,
                                     const gd_adam* opt, float* weights_dev, float* exp_avg_dev, float* exp_avg_sq_dev,
                                     void* stream) {
    GD_CHECK_ARG(peer_ptrs_host && src_dev && err_dev && opt && weights_dev && exp_avg_dev && exp_avg_sq_dev,
                 "gd_p2p_allreduce_adam: NULL argument");
    GD_CHECK_ARG(world >= 1 && world <= gd::kMaxWorld && rank >= 0 && rank < world, "gd_p2p_allreduce_adam: bad rank %d / world %d",
                 rank, world);
    GD_CHECK_ARG(n > 0 && epoch > 0 && opt->step >= 1, "gd_p2p_allreduce_adam: n, epoch and step must be positive");
    GD_CHECK_ARG(opt->beta1 >= 0. && opt->beta1 < 1. && opt->beta2 >= 0. && opt->beta2 < 1. && opt->eps >= 0. && opt->lr >= 0.
                 && opt->weight_decay >= 0., "gd_p2p_allreduce_adam: invalid hyper-parameters");
    gd::P2PParams p;
    for (int r = 0; r < world; ++r) p.peer[r] = reinterpret_cast<float*>((uintptr_t)peer_ptrs_host[r]);
    p.src = src_dev; p.dst = dst_dev; p.err = err_dev; p.world = world; p.rank = rank; p.n = n; p.n_pad = (n + 31) / 32 * 32;
    p.epoch = epoch; p.scale = scale;
    p.w = weights_dev; p.m = exp_avg_dev; p.v = exp_avg_sq_dev;
    p.adam = gd::adam_coef(opt->lr, opt->beta1, opt->beta2, opt->eps, opt->weight_decay, opt->step);
    gd::p2p_allreduce_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(p);
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

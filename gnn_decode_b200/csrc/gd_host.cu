// gd_decode_host: the end-to-end call with HOST buffers.  Replaces the reference's
// `datas = datas.to(device); pred = decoder(datas)` (quantum/decoder_v2_4.py:332-334) plus the
// read-back of the prediction.  The batch is cut into chunks that flow through two streams
// (H2D copy -> fused decode kernel -> D2H copy), so copies of one chunk overlap the kernel of
// the other and the second kernel's CTAs back-fill the SMs the first one's last wave leaves idle.
#include "gd_common.cuh"
#include <algorithm>

namespace gd {

struct HostCtx {
    cudaStream_t st[2] = {nullptr, nullptr};
    float* x_dev = nullptr;
    float* prob_dev = nullptr;
    uint8_t* hard_dev = nullptr;
    float* w_dev = nullptr;
    int64_t cap_B = 0;
    int64_t cap_w = 0;
};

static cudaError_t ensure(HostCtx* c, const gd_graph* g, int64_t B, int64_t n_w) {
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i)
        if (!c->st[i]) e = cudaStreamCreateWithFlags(&c->st[i], cudaStreamNonBlocking);
    if (e == cudaSuccess && B > c->cap_B) {
        if (c->x_dev) cudaFree(c->x_dev);
        if (c->prob_dev) cudaFree(c->prob_dev);
        if (c->hard_dev) cudaFree(c->hard_dev);
        c->x_dev = nullptr; c->prob_dev = nullptr; c->hard_dev = nullptr; c->cap_B = 0;
        e = cudaMalloc((void**)&c->x_dev, (size_t)B * g->N * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc((void**)&c->prob_dev, (size_t)B * g->V * sizeof(float));
        if (e == cudaSuccess) e = cudaMalloc((void**)&c->hard_dev, (size_t)B * g->V);
        if (e == cudaSuccess) c->cap_B = B;
    }
    if (e == cudaSuccess && n_w > c->cap_w) {
        if (c->w_dev) cudaFree(c->w_dev);
        c->w_dev = nullptr; c->cap_w = 0;
        e = cudaMalloc((void**)&c->w_dev, (size_t)n_w * sizeof(float));
        if (e == cudaSuccess) c->cap_w = n_w;
    }
    return e;
}

}  // namespace gd

void gd_host_ctx_destroy(gd_graph* g) {
    gd::HostCtx* c = static_cast<gd::HostCtx*>(g->host_ctx);
    if (!c) return;
    for (int i = 0; i < 2; ++i)
        if (c->st[i]) cudaStreamDestroy(c->st[i]);
    if (c->x_dev) cudaFree(c->x_dev);
    if (c->prob_dev) cudaFree(c->prob_dev);
    if (c->hard_dev) cudaFree(c->hard_dev);
    if (c->w_dev) cudaFree(c->w_dev);
    delete c;
    g->host_ctx = nullptr;
}

extern "C" int gd_decode_host(const gd_graph* gc, const gd_model* model, const float* weights_host,
                              const float* x_host, float* prob_host, uint8_t* hard_host, int64_t B) {
    gd_graph* g = const_cast<gd_graph*>(gc);
    GD_CHECK_ARG(g != nullptr, "gd_decode_host: graph is NULL");
    GD_CHECK_ARG(gd_model_valid(model), "gd_decode_host: invalid model");
    GD_CHECK_ARG(B >= 0 && B < ((int64_t)1 << 31), "gd_decode_host: B out of range");
    GD_CHECK_ARG(model->flags == 0, "gd_decode_host: per-iteration outputs (GD_FLAG_ALL_ITERS) are device-path only");
    if (B == 0) return GD_OK;
    GD_CHECK_ARG(x_host != nullptr, "gd_decode_host: x is NULL");
    GD_CHECK_ARG(prob_host || hard_host, "gd_decode_host: no output requested");
    const int64_t n_w = gd_weights_size(model);
    GD_CHECK_ARG(n_w == 0 || weights_host, "gd_decode_host: weights is NULL");

    int prev = 0;
    GD_CUDA(cudaGetDevice(&prev));
    if (prev != g->device) GD_CUDA(cudaSetDevice(g->device));
    std::unique_lock<std::mutex> lk(g->mu);
    if (!g->host_ctx) g->host_ctx = new gd::HostCtx();
    gd::HostCtx* c = static_cast<gd::HostCtx*>(g->host_ctx);
    cudaError_t e = gd::ensure(c, g, B, n_w);
    lk.unlock();  // gd_decode_fwd takes the same mutex for the streamed workspace
    int rc = GD_OK;
    if (e == cudaSuccess && n_w) {
        e = cudaMemcpyAsync(c->w_dev, weights_host, (size_t)n_w * sizeof(float), cudaMemcpyHostToDevice, c->st[0]);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->st[0]);
    }
    if (e == cudaSuccess) {
        // chunks: multiples of 8 syndromes (output alignment); >= 4096 syndromes each, at most 4
        int n_chunks = (int)std::min<int64_t>(4, std::max<int64_t>(1, B / 8192));
        gd_launch_info li;
        if (gd_decode_launch_info(g, model, B, &li) == GD_OK && !li.resident) n_chunks = 1;  // one shared slab
        int64_t per = ((B + n_chunks - 1) / n_chunks + 7) / 8 * 8;
        for (int k = 0; k < n_chunks && e == cudaSuccess && rc == GD_OK; ++k) {
            const int64_t b0 = (int64_t)k * per;
            const int64_t nb = std::min<int64_t>(per, B - b0);
            if (nb <= 0) break;
            cudaStream_t st = c->st[k & 1];
            e = cudaMemcpyAsync(c->x_dev + b0 * g->N, x_host + b0 * g->N, (size_t)nb * g->N * sizeof(float),
                                cudaMemcpyHostToDevice, st);
            if (e != cudaSuccess) break;
            rc = gd_decode_fwd(g, model, c->w_dev, c->x_dev + b0 * g->N, prob_host ? c->prob_dev + b0 * g->V : nullptr,
                               nullptr, hard_host ? c->hard_dev + b0 * g->V : nullptr, nb, st);
            if (rc != GD_OK) break;
            if (prob_host)
                e = cudaMemcpyAsync(prob_host + b0 * g->V, c->prob_dev + b0 * g->V, (size_t)nb * g->V * sizeof(float),
                                    cudaMemcpyDeviceToHost, st);
            if (e == cudaSuccess && hard_host)
                e = cudaMemcpyAsync(hard_host + b0 * g->V, c->hard_dev + b0 * g->V, (size_t)nb * g->V,
                                    cudaMemcpyDeviceToHost, st);
        }
        cudaError_t s0 = cudaStreamSynchronize(c->st[0]);
        cudaError_t s1 = cudaStreamSynchronize(c->st[1]);
        if (e == cudaSuccess) e = s0 != cudaSuccess ? s0 : s1;
    }
    if (prev != g->device) cudaSetDevice(prev);
    if (rc != GD_OK) return rc;
    GD_CUDA(e);
    return GD_OK;
}

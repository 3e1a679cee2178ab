// Internal interface between the resident (gd_decode.cu) and streamed (gd_streamed.cu) decoders.
#pragma once
#include "gd_common.cuh"

namespace gd {

int streamed_launch_info(const gd_graph* g, const gd_model* model, int64_t B, gd_launch_info* out);
int streamed_decode(gd_graph* g, const gd_model* model, const float* weights_dev, const float* x_dev,
                    float* prob_dev, float* logit_dev, uint8_t* hard_dev, int64_t B, cudaStream_t st);

bool streamed_tma_launch_info(const gd_graph* g, const gd_model* model, int64_t B, gd_launch_info* out);
int streamed_tma_decode(gd_graph* g, const gd_model* model, const float* weights_dev, const float* x_dev,
                        float* prob_dev, float* logit_dev, uint8_t* hard_dev, int64_t B, cudaStream_t st);

bool light_launch_info(const gd_graph* g, const gd_model* model, int64_t B, gd_launch_info* out);
int light_decode(const gd_graph* g, const gd_model* model, const float* weights_dev, const float* x_dev,
                 float* prob_dev, float* logit_dev, uint8_t* hard_dev, int64_t B, cudaStream_t st);

}  // namespace gd

// Network handle shared by the fp32 and tcgen05 back ends.
#pragma once
#include <map>
#include <string>
#include <vector>
#include "common.cuh"

struct dxi_net {
  int kind = 0;
  dxi_net_cfg cfg{};
  int device = 0;
  bool finalized = false;
  std::map<std::string, std::vector<float>> host;        // checkpoint tensors by Keras name
  std::map<std::string, std::vector<int64_t>> shapes;
  // fp32 device copies (one arena) -- used by the fp32 back end and by the CUDA-core layers
  float* d_arena = nullptr;
  std::map<std::string, size_t> d_offset;                 // in floats
  // tcgen05 back end: packed fp16 hi/lo operand images (see tcn_umma.cu)
  void* d_umma = nullptr;
  size_t umma_bytes = 0;
  std::vector<size_t> umma_stage_offset;
  // depth-first ResNetV2 path (tcn_chain.cu): plain fp16 hi/lo weight matrices + aux records, and the tensor maps over them
  void* d_chain = nullptr;
  size_t chain_aux_offset = 0;
  size_t chain_stem_aux_offset = 0, chain_head_aux_offset = 0;
  alignas(64) unsigned char chain_tm[5][128] = {};      // W1, W2, W3 of the blocks; first layer; output layer

  const float* dev_tensor(int layer, const char* var) const {
    char name[96];
    snprintf(name, sizeof(name), "layer_with_weights-%d/%s", layer, var);
    auto it = d_offset.find(name);
    return it == d_offset.end() ? nullptr : d_arena + it->second;
  }
  const std::vector<float>* host_tensor(int layer, const char* var) const {
    char name[96];
    snprintf(name, sizeof(name), "layer_with_weights-%d/%s", layer, var);
    auto it = host.find(name);
    return it == host.end() ? nullptr : &it->second;
  }
};

namespace dxi {
int64_t resnet_f32_workspace_bytes(const dxi_net& net, int B, int T);
int resnet_f32_forward(const dxi_net& net, const float* mag, int B, int T, float* xbar, void* ws, size_t ws_bytes,
                       cudaStream_t st);
int64_t resnet_umma_workspace_bytes(const dxi_net& net, int B, int T);
int resnet_umma_prepare(dxi_net& net, cudaStream_t st);
int resnet_umma_forward(const dxi_net& net, const float* mag, int B, int T, float* xbar, void* ws, size_t ws_bytes,
                        cudaStream_t st);
bool resnet_chain_supported(const dxi_net& net);
float resnet_first_operand_scale(const dxi_net& net);
float weight_pow2_scale(const float* W, size_t n, bool half);
int resnet_chain_prepare(dxi_net& net, cudaStream_t st);
size_t resnet_chain_extra_workspace(const dxi_net& net, int B, int tiles);
int resnet_chain_blocks(const dxi_net& net, float* h, const float2* stem_stats, int B, int T, void* extra, int n_sm, cudaStream_t st);
bool resnet_chain_fused(const dxi_net& net);
int resnet_chain_network(const dxi_net& net, const float* mag, float* xbar, int B, int T, void* extra, int n_sm, cudaStream_t st);
int64_t mhanet_workspace_bytes(const dxi_net& net, int B, int T);
int mhanet_forward(const dxi_net& net, const float* mag, int B, int T, float* xbar, void* ws, size_t ws_bytes,
                   cudaStream_t st);
int mhanet_umma_prepare(dxi_net& net, cudaStream_t st);
size_t mhanet_umma_attention_workspace(const dxi_net& net, int B, int T);
int mhanet_umma_qkv(const dxi_net& net, int blk, const float* x, int B, int T, float* qkv, void* kv, cudaStream_t st);
int mhanet_umma_attention(const dxi_net& net, const float* qkv, const uint8_t* valid, int B, int T, float* att, void* kv, bool kv_packed, cudaStream_t st);
int mhanet_umma_linear(const dxi_net& net, int image, int epi, const float* A, int lda, const float* bias, const float* res,
                       const float* gamma, const float* beta, const float* pos, int T, float* out, int ldo, int M, int Nr, int Kr,
                       cudaStream_t st);
}  // namespace dxi

// dxi_net_* : network handle management (weights by checkpoint name, precision selection, dispatch).
#include "net.cuh"

using namespace dxi;

namespace {

// Expected tensor shapes, in the checkpoint's layer order (see deepxi_b200/weights.py and
// SURVEY appendix A): ResNetV2 tcn.py:116-225, MHANetV3 attention.py:387-442.
std::map<std::string, std::vector<int64_t>> expected_shapes(int kind, const dxi_net_cfg& c) {
  std::map<std::string, std::vector<int64_t>> s;
  auto nm = [](int li, const char* v) { char b[96]; snprintf(b, sizeof(b), "layer_with_weights-%d/%s", li, v); return std::string(b); };
  if (kind == DXI_NET_RESNETV2) {
    s[nm(0, "kernel")] = {1, c.n_feat, c.d_model};
    s[nm(0, "bias")] = {c.d_model};
    s[nm(1, "gamma")] = {c.d_model};
    int li = 2;
    for (int i = 0; i < c.n_blocks; ++i) {
      s[nm(li, "kernel")] = {1, c.d_model, c.d_f};      s[nm(li, "bias")] = {c.d_f};
      s[nm(li + 1, "kernel")] = {c.k, c.d_f, c.d_f};    s[nm(li + 1, "bias")] = {c.d_f};
      s[nm(li + 2, "kernel")] = {1, c.d_f, c.d_model};  s[nm(li + 2, "bias")] = {c.d_model};
      li += 3;
    }
    s[nm(li, "kernel")] = {1, c.d_model, c.n_outp};
    s[nm(li, "bias")] = {c.n_outp};
  } else if (kind == DXI_NET_RESNETV3) {      // ResNetV2 without the LayerNorm weights of the first layer (tcn.py:241-244)
    s[nm(0, "kernel")] = {1, c.n_feat, c.d_model};
    s[nm(0, "bias")] = {c.d_model};
    int li = 1;
    for (int i = 0; i < c.n_blocks; ++i) {
      s[nm(li, "kernel")] = {1, c.d_model, c.d_f};      s[nm(li, "bias")] = {c.d_f};
      s[nm(li + 1, "kernel")] = {c.k, c.d_f, c.d_f};    s[nm(li + 1, "bias")] = {c.d_f};
      s[nm(li + 2, "kernel")] = {1, c.d_f, c.d_model};  s[nm(li + 2, "bias")] = {c.d_model};
      li += 3;
    }
    s[nm(li, "kernel")] = {1, c.d_model, c.n_outp};
    s[nm(li, "bias")] = {c.n_outp};
  } else if (kind == DXI_NET_RESNET) {        // tcn.py:17-114; order and sizes as in log/summary/resnet-1.0c.txt (1 975 553 parameters)
    s[nm(0, "kernel")] = {1, c.n_feat, c.d_model};
    s[nm(1, "gamma")] = {c.d_model};  s[nm(1, "beta")] = {c.d_model};
    int li = 2;
    for (int i = 0; i < c.n_blocks; ++i) {
      s[nm(li, "gamma")] = {c.d_model};    s[nm(li, "beta")] = {c.d_model};      s[nm(li + 1, "kernel")] = {1, c.d_model, c.d_f};
      s[nm(li + 2, "gamma")] = {c.d_f};    s[nm(li + 2, "beta")] = {c.d_f};      s[nm(li + 3, "kernel")] = {c.k, c.d_f, c.d_f};
      s[nm(li + 4, "gamma")] = {c.d_f};    s[nm(li + 4, "beta")] = {c.d_f};      s[nm(li + 5, "kernel")] = {1, c.d_f, c.d_model};
      s[nm(li + 5, "bias")] = {c.d_model};
      li += 6;
    }
    s[nm(li, "kernel")] = {1, c.d_model, c.n_outp};
    s[nm(li, "bias")] = {c.n_outp};
  } else {
    const int64_t dk = c.d_model / c.n_heads, dff = 4 * c.d_model;
    s[nm(0, "kernel")] = {1, c.n_feat, c.d_model};
    s[nm(1, "gamma")] = {c.d_model};  s[nm(1, "beta")] = {c.d_model};
    s[nm(2, "embeddings")] = {c.max_len, c.d_model};
    int li = 3;
    for (int i = 0; i < c.n_blocks; ++i) {
      s[nm(li, "query_kernel")] = {c.n_heads, c.d_model, dk};
      s[nm(li, "key_kernel")] = {c.n_heads, c.d_model, dk};
      s[nm(li, "value_kernel")] = {c.n_heads, c.d_model, dk};
      s[nm(li, "projection_kernel")] = {c.n_heads, dk, c.d_model};
      s[nm(li + 1, "gamma")] = {c.d_model};  s[nm(li + 1, "beta")] = {c.d_model};
      s[nm(li + 2, "kernel")] = {1, c.d_model, dff};  s[nm(li + 2, "bias")] = {dff};
      s[nm(li + 3, "kernel")] = {1, dff, c.d_model};  s[nm(li + 3, "bias")] = {c.d_model};
      s[nm(li + 4, "gamma")] = {c.d_model};  s[nm(li + 4, "beta")] = {c.d_model};
      li += 5;
    }
    s[nm(li, "kernel")] = {1, c.d_model, c.n_outp};
    s[nm(li, "bias")] = {c.n_outp};
  }
  return s;
}

}  // namespace

extern "C" DXI_API int dxi_net_create(dxi_net_t** h, int kind, const dxi_net_cfg* cfg) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(h && cfg, "dxi_net_create: null argument");
  if (kind < DXI_NET_RESNETV2 || kind > DXI_NET_RESNETV3) { set_error("Invalid network type."); return DXI_E_INVALID; }
  if ((kind == DXI_NET_RESNET || kind == DXI_NET_RESNETV3) && cfg->precision != DXI_PREC_F32) {
    set_error("ResNet / ResNetV3 are built for precision f32 only"); return DXI_E_INVALID;
  }
  DXI_REQUIRE(cfg->precision >= DXI_PREC_F32 && cfg->precision <= DXI_PREC_F16, "dxi_net_create: bad precision");
  DXI_REQUIRE(cfg->n_feat > 0 && cfg->n_outp > 0 && cfg->d_model > 0 && cfg->n_blocks > 0, "dxi_net_create: bad sizes");
  if (kind != DXI_NET_MHANETV3) {
    DXI_REQUIRE(cfg->padding == DXI_PAD_CAUSAL || cfg->padding == DXI_PAD_SAME, "dxi_net_create: bad padding");
    DXI_REQUIRE(cfg->max_d_rate >= 1 && (cfg->max_d_rate & (cfg->max_d_rate - 1)) == 0, "dxi_net_create: max_d_rate must be a power of two");
  } else {
    DXI_REQUIRE(cfg->n_heads > 0 && cfg->d_model % cfg->n_heads == 0 && cfg->max_len > 0, "dxi_net_create: bad attention sizes");
    DXI_REQUIRE(cfg->mask_mode == DXI_MASK_NONE || cfg->mask_mode == DXI_MASK_CAUSAL_PAD, "dxi_net_create: bad mask_mode");
  }
  dxi_net* n = new dxi_net();
  n->kind = kind;
  n->cfg = *cfg;
  cudaGetDevice(&n->device);
  *h = n;
  return DXI_OK;
}

extern "C" DXI_API int dxi_net_load(dxi_net_t* h, const char* tensor_name, const float* host_data, const int64_t* shape, int rank) {
  DXI_REQUIRE(h && tensor_name && host_data && shape && rank > 0 && rank <= 4, "dxi_net_load: bad argument");
  auto exp = expected_shapes(h->kind, h->cfg);
  auto it = exp.find(tensor_name);
  if (it == exp.end()) { set_error("dxi_net_load: unexpected tensor '%s' for this network", tensor_name); return DXI_E_INVALID; }
  std::vector<int64_t> shp(shape, shape + rank);
  if (shp != it->second) { set_error("dxi_net_load: shape mismatch for '%s'", tensor_name); return DXI_E_INVALID; }
  int64_t n = 1;
  for (auto d : shp) n *= d;
  h->host[tensor_name].assign(host_data, host_data + n);
  h->shapes[tensor_name] = shp;
  h->finalized = false;
  return DXI_OK;
}

extern "C" DXI_API int dxi_net_finalize(dxi_net_t* h, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(h, "dxi_net_finalize: null handle");
  {
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev != h->device) { set_error("dxi_net_finalize: handle belongs to device %d, current device is %d", h->device, dev); return DXI_E_STATE; }
  }
  auto exp = expected_shapes(h->kind, h->cfg);
  for (auto& kv : exp)
    if (!h->host.count(kv.first)) { set_error("dxi_net_finalize: tensor '%s' was not loaded", kv.first.c_str()); return DXI_E_STATE; }
  cudaStream_t st = as_stream(stream);
  if (h->kind == DXI_NET_MHANETV3) {
    // fused [d_model][3*d_model] projection per block: column = which*d_model + head*d_k + o
    // (query/key/value_kernel are [heads][d_model][d_k], the tfa einsum "...NI,HIO->...NHO")
    const int d = h->cfg.d_model, H = h->cfg.n_heads, dk = d / H;
    for (int blk = 0; blk < h->cfg.n_blocks; ++blk) {
      const int li = 3 + 5 * blk;
      char nm[64];
      snprintf(nm, sizeof(nm), "packed-%d/qkv", li);
      std::vector<float> packed((size_t)d * 3 * d);
      const char* names[3] = {"query_kernel", "key_kernel", "value_kernel"};
      for (int w = 0; w < 3; ++w) {
        const std::vector<float>& src = *h->host_tensor(li, names[w]);
        for (int hh = 0; hh < H; ++hh)
          for (int i = 0; i < d; ++i)
            for (int o = 0; o < dk; ++o) packed[(size_t)i * 3 * d + w * d + hh * dk + o] = src[((size_t)hh * d + i) * dk + o];
      }
      h->host[nm] = std::move(packed);
    }
  }
  // fp32 arena (every tensor 256-byte aligned)
  size_t total = 0;
  h->d_offset.clear();
  for (auto& kv : h->host) { h->d_offset[kv.first] = total; total += (kv.second.size() + 63) / 64 * 64; }
  if (h->d_arena) { cudaFree(h->d_arena); h->d_arena = nullptr; }
  DXI_CUDA(cudaMalloc(&h->d_arena, total * sizeof(float)));
  for (auto& kv : h->host)
    DXI_CUDA(cudaMemcpyAsync(h->d_arena + h->d_offset[kv.first], kv.second.data(), kv.second.size() * sizeof(float),
                             cudaMemcpyHostToDevice, st));
  if (h->kind == DXI_NET_RESNETV2 && h->cfg.precision != DXI_PREC_F32) {
    if (int rc = resnet_umma_prepare(*h, st)) return rc;
  }
  if (h->kind == DXI_NET_MHANETV3 && h->cfg.precision != DXI_PREC_F32) {
    if (int rc = mhanet_umma_prepare(*h, st)) return rc;
  }
  DXI_CUDA(cudaStreamSynchronize(st));
  h->finalized = true;
  return DXI_OK;
}

extern "C" DXI_API int64_t dxi_net_workspace_bytes(const dxi_net_t* h, int B, int Tmax) {
  if (!h || B < 0 || Tmax < 0) return DXI_E_INVALID;
  if (h->kind != DXI_NET_MHANETV3)
    return h->cfg.precision == DXI_PREC_F32 ? resnet_f32_workspace_bytes(*h, B, Tmax) : resnet_umma_workspace_bytes(*h, B, Tmax);
  return mhanet_workspace_bytes(*h, B, Tmax);
}

extern "C" DXI_API int dxi_net_forward(dxi_net_t* h, const float* mag, int B, int Tmax, float* xbar, void* workspace,
                               size_t workspace_bytes, void* stream) {
  if (int rc = check_device()) return rc;
  DXI_REQUIRE(h && mag && xbar, "dxi_net_forward: null argument");
  DXI_REQUIRE(B >= 0 && Tmax >= 0, "dxi_net_forward: bad shape");
  if (!h->finalized) { set_error("dxi_net_forward: call dxi_net_finalize first"); return DXI_E_STATE; }
  {   // the handle's weights live on the device that was current at creation: refuse anything else instead of faulting
    int dev = -1;
    cudaGetDevice(&dev);
    if (dev != h->device) { set_error("dxi_net_forward: handle belongs to device %d, current device is %d", h->device, dev); return DXI_E_STATE; }
  }
  if (B == 0 || Tmax == 0) return DXI_OK;
  DXI_REQUIRE(workspace, "dxi_net_forward: null workspace");
  cudaStream_t st = as_stream(stream);
  if (h->kind != DXI_NET_MHANETV3) {
    if (h->cfg.precision == DXI_PREC_F32) return resnet_f32_forward(*h, mag, B, Tmax, xbar, workspace, workspace_bytes, st);
    return resnet_umma_forward(*h, mag, B, Tmax, xbar, workspace, workspace_bytes, st);
  }
  return mhanet_forward(*h, mag, B, Tmax, xbar, workspace, workspace_bytes, st);
}

extern "C" DXI_API int dxi_net_destroy(dxi_net_t* h) {
  if (!h) return DXI_OK;
  if (h->d_arena) cudaFree(h->d_arena);
  if (h->d_umma) cudaFree(h->d_umma);
  if (h->d_chain) cudaFree(h->d_chain);
  delete h;
  return DXI_OK;
}

// 20 m -> 10 m super-resolution CNN (superresolve_graph.pb; reference call site
// src/download_and_predict_job.py:110-120, sess.run(superresolve_logits, ...)).
// pb: [MirrorPad(1) + Conv2D 3x3 VALID + BiasAdd] x6; Relu after in/01/11; x0.1 residual adds
// (pb:Mul, pb:Add, pb:Mul_1, pb:Add_1); Tanh; + Placeholder_1 (bilinear) = pb:Add_2.
#include "stc_common.cuh"
#include <cstring>

int pack_and_upload_conv(stc_ctx* ctx, const float* w, int Cin, int Cout, const std::vector<int>& chanmap, int Npad, uint4** dptr);

// activation arena of one (N, H, W) launch shape; superresolve_large_tile alternates between two shapes per tile (all
// windows but one, then the overlapping corner window), so the last few plans are kept instead of cudaFree / cudaMalloc /
// memset of a 2.5 GB arena on every switch (that was 5-15 ms of the 6-23 ms the stage took)
struct SrPlan {
  void* arena = nullptr; size_t arena_bytes = 0;
  int N = 0, H = 0, W = 0;
  Act X, A, Bf;               // 10->16 ch input (2 chunks), two 32-ch ping-pong activations
  Raw raw, skip;              // conv output (32 ch), fp32 skip path (32 ch)
  uint64_t last_use = 0;
};
struct SrState : SrPlan {
  uint4* w[6] = {nullptr};
  float* bias = nullptr;      // 6 x 32 floats (out layer padded to 16)
  bool ready = false;
  std::vector<SrPlan> cache;  // plans not in use (the current one lives in the base class)
  uint64_t tick = 0;
};

__device__ __forceinline__ uint4 sr_pack8(const float* v) {
  uint4 r;
  __half2 h0 = __floats2half2_rn(v[0], v[1]), h1 = __floats2half2_rn(v[2], v[3]);
  __half2 h2 = __floats2half2_rn(v[4], v[5]), h3 = __floats2half2_rn(v[6], v[7]);
  r.x = *reinterpret_cast<uint32_t*>(&h0); r.y = *reinterpret_cast<uint32_t*>(&h1);
  r.z = *reinterpret_cast<uint32_t*>(&h2); r.w = *reinterpret_cast<uint32_t*>(&h3);
  return r;
}

__device__ __forceinline__ void sr_write_reflect(uint4* base, int64_t plane, int chunks, int b, int yp, int xp, int Hp, int Wp,
                                                 const uint4* vals) {
  int ys[3] = {yp, (yp == 2) ? 0 : -1, (yp == Hp - 3) ? Hp - 1 : -1};
  int xs[3] = {xp, (xp == 2) ? 0 : -1, (xp == Wp - 3) ? Wp - 1 : -1};
  for (int i = 0; i < 3; ++i) {
    if (ys[i] < 0) continue;
    for (int j = 0; j < 3; ++j) {
      if (xs[j] < 0) continue;
      int64_t P = ((int64_t)b * Hp + ys[i]) * Wp + xs[j];
      for (int c = 0; c < chunks; ++c) base[(int64_t)c * plane + P] = vals[c];
    }
  }
}

// x [N,H,W,10] f32 -> fp16 16-channel activation with reflect border
__global__ void __launch_bounds__(256) sr_prep_kernel(const float* __restrict__ x, int N, int H, int W, uint4* dst, int64_t plane) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)N * H * W) return;
  int xw = (int)(idx % W); int64_t r = idx / W; int y = (int)(r % H); int b = (int)(r / H);
  float v[16];
#pragma unroll
  for (int c = 0; c < 16; ++c) v[c] = 0.f;
  const float* s = x + idx * 10;
#pragma unroll
  for (int c = 0; c < 10; ++c) v[c] = s[c];
  uint4 o[2] = {sr_pack8(v), sr_pack8(v + 8)};
  sr_write_reflect(dst, plane, 2, b, y + 1, xw + 1, H + 2, W + 2, o);
}

// mode 0: act = fp16(raw)                       (after Relu layers; optionally also skip = raw)
// mode 1: skip += 0.1*raw; act = fp16(skip)     (residual join)
// mode 2: out = tanh(raw[0:6]) + bilinear       (final)
struct SrApply {
  const float4* raw; int64_t raw_plane;
  float4* skip; int64_t skip_plane; int write_skip;
  uint4* dst; int64_t dst_plane;
  const float* bil; int bil_stride, bil_off; float* out;   // bilinear input: [.., bil_stride] floats per pixel, band k at bil_off + k
  int N, H, W; int mode;
};

__global__ void __launch_bounds__(256) sr_apply_kernel(SrApply p) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)p.N * p.H * p.W) return;
  int xw = (int)(idx % p.W); int64_t r = idx / p.W; int y = (int)(r % p.H); int b = (int)(r / p.H);
  int Hp = p.H + 2, Wp = p.W + 2;
  int64_t P = ((int64_t)b * Hp + y + 1) * Wp + xw + 1;
  if (p.mode == 2) {
    float4 a = p.raw[P], c = p.raw[p.raw_plane + P];
    float t[6] = {a.x, a.y, a.z, a.w, c.x, c.y};
#pragma unroll
    for (int k = 0; k < 6; ++k) p.out[idx * 6 + k] = tanhf(t[k]) + p.bil[idx * p.bil_stride + p.bil_off + k];
    return;
  }
  uint4 o[4];
#pragma unroll
  for (int c8 = 0; c8 < 4; ++c8) {
    float v[8];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      float4 a = p.raw[(int64_t)(2 * c8 + h) * p.raw_plane + P];
      if (p.mode == 1) {
        float4 s = p.skip[(int64_t)(2 * c8 + h) * p.skip_plane + P];
        a = make_float4(s.x + 0.1f * a.x, s.y + 0.1f * a.y, s.z + 0.1f * a.z, s.w + 0.1f * a.w);
        p.skip[(int64_t)(2 * c8 + h) * p.skip_plane + P] = a;
      } else if (p.write_skip) {
        p.skip[(int64_t)(2 * c8 + h) * p.skip_plane + P] = a;
      }
      v[4 * h] = a.x; v[4 * h + 1] = a.y; v[4 * h + 2] = a.z; v[4 * h + 3] = a.w;
    }
    o[c8] = sr_pack8(v);
  }
  sr_write_reflect(p.dst, p.dst_plane, 4, b, y + 1, xw + 1, Hp, Wp, o);
}

static const char* SRN[6] = {"in", "r01", "r02", "r11", "r12", "out"};

int sr_finalize_weights(stc_ctx* ctx) {
  SrState* s = (SrState*)ctx->sr;
  if (!s) { s = new SrState(); ctx->sr = s; }
  std::vector<float> bias(6 * 32, 0.f);
  for (int i = 0; i < 6; ++i) {
    int cin = (i == 0) ? 10 : 32, cout = (i == 5) ? 6 : 32, npad = (i == 5) ? 16 : 32;
    auto wi = ctx->host_w.find(std::string("sr.") + SRN[i] + ".w");
    auto bi = ctx->host_w.find(std::string("sr.") + SRN[i] + ".b");
    if (wi == ctx->host_w.end() || bi == ctx->host_w.end() || wi->second.size() != (size_t)9 * cin * cout || bi->second.size() != (size_t)cout)
      STC_FAIL(STC_ERR_STATE, std::string("missing super-resolve weights for layer ") + SRN[i]);
    int cpad = (i == 0) ? 16 : 32;
    std::vector<int> map(cpad, -1);
    for (int c = 0; c < cin; ++c) map[c] = c;
    int rc = pack_and_upload_conv(ctx, wi->second.data(), cin, cout, map, npad, &s->w[i]); if (rc) return rc;
    for (int c = 0; c < cout; ++c) bias[i * 32 + c] = bi->second[c];
  }
  if (s->bias) cudaFree(s->bias);
  STC_CUDA(cudaMalloc((void**)&s->bias, bias.size() * sizeof(float)));
  STC_CUDA(cudaMemcpy(s->bias, bias.data(), bias.size() * sizeof(float), cudaMemcpyHostToDevice));
  s->ready = true;
  return STC_OK;
}

void sr_destroy(stc_ctx* ctx) {
  SrState* s = (SrState*)ctx->sr;
  if (!s) return;
  for (int i = 0; i < 6; ++i) cudaFree(s->w[i]);
  cudaFree(s->bias); cudaFree(s->arena);
  for (auto& pl : s->cache) cudaFree(pl.arena);
  delete s; ctx->sr = nullptr;
}

static int sr_plan(stc_ctx* ctx, SrState* s, int N, int H, int W) {
  s->last_use = ++s->tick;
  if (s->arena && s->N == N && s->H == H && s->W == W) return STC_OK;
  if (s->arena) { s->cache.push_back(*static_cast<SrPlan*>(s)); s->arena = nullptr; }
  for (size_t i = 0; i < s->cache.size(); ++i)
    if (s->cache[i].N == N && s->cache[i].H == H && s->cache[i].W == W) {
      *static_cast<SrPlan*>(s) = s->cache[i]; s->cache.erase(s->cache.begin() + i); s->last_use = s->tick;
      return STC_OK;
    }
  while (s->cache.size() > 3) {                      // keep at most 3 idle plans: drop the least recently used
    size_t lru = 0;
    for (size_t i = 1; i < s->cache.size(); ++i) if (s->cache[i].last_use < s->cache[lru].last_use) lru = i;
    STC_CUDA(cudaStreamSynchronize(ctx->stream));
    cudaFree(s->cache[lru].arena); s->cache.erase(s->cache.begin() + lru);
  }
  int Hp = H + 2, Wp = W + 2;
  int64_t P = (int64_t)N * Hp * Wp;
  int guard = ((Wp + 2 + 544 + 7) / 8) * 8;
  int64_t plane = P + 2 * guard;
  int64_t rp = (P + 511) / 512 * 512;
  size_t bytes = (size_t)(2 + 4 + 4) * plane * 16 + (size_t)(8 + 8) * rp * 16;
  STC_CUDA(cudaMalloc(&s->arena, bytes));
  STC_CUDA(cudaMemsetAsync(s->arena, 0, bytes, ctx->stream));
  char* p = (char*)s->arena;
  auto mk = [&](Act& a, int chunks) {
    a.base = (uint4*)p; a.plane = plane; a.guard = guard; a.chunks = chunks; a.B = N; a.Hp = Hp; a.Wp = Wp;
    p += (size_t)chunks * plane * 16;
  };
  mk(s->X, 2); mk(s->A, 4); mk(s->Bf, 4);
  s->raw.base = (float4*)p; s->raw.plane = rp; s->raw.N = 32; p += (size_t)8 * rp * 16;
  s->skip.base = (float4*)p; s->skip.plane = rp; s->skip.N = 32;
  s->N = N; s->H = H; s->W = W; s->arena_bytes = bytes;
  return STC_OK;
}

// dst / skip_mode / out: the fused epilogues (ConvParams); all null = raw fp32 output for sr_apply_kernel
static int sr_conv(stc_ctx* ctx, SrState* s, int layer, const Act& in, int mode, Act* dst = nullptr, int skip_mode = 0,
                   float* out = nullptr, const float* bil = nullptr, int bil_stride = 0, int bil_off = 0) {
  ConvParams cp; memset(&cp, 0, sizeof(cp));
  cp.a0[0] = in.at(0); cp.a0_plane = in.plane; cp.k0steps = in.chunks / 2;
  cp.w[0] = s->w[layer]; cp.out[0] = s->raw.base; cp.out_plane = s->raw.plane;
  cp.bias = s->bias + layer * 32;
  cp.N = (layer == 5) ? 16 : 32; cp.G = 1;
  cp.B = in.B; cp.Hp = in.Hp; cp.Wp = in.Wp; cp.Ptot = in.Ptot();
  cp.vy0 = 1; cp.vy1 = in.Hp - 1; cp.vx0 = 1; cp.vx1 = in.Wp - 1;
  cp.mode = mode;
  if (dst) { cp.act16 = dst->at(0); cp.act16_plane = dst->plane; cp.skip = s->skip.base; cp.skip_plane = s->skip.plane; cp.skip_mode = skip_mode; }
  if (out) { cp.sr_out = out; cp.sr_bil = bil; cp.sr_bil_stride = bil_stride; cp.sr_bil_off = bil_off; }
  return launch_conv(ctx, cp, 1);
}

static int sr_apply(stc_ctx* ctx, SrState* s, int mode, int write_skip, Act* dst, const float* bil, float* out,
                    int bil_stride = 6, int bil_off = 0) {
  SrApply ap; memset(&ap, 0, sizeof(ap));
  ap.raw = s->raw.base; ap.raw_plane = s->raw.plane; ap.skip = s->skip.base; ap.skip_plane = s->skip.plane;
  ap.write_skip = write_skip;
  if (dst) { ap.dst = dst->at(0); ap.dst_plane = dst->plane; }
  ap.bil = bil; ap.bil_stride = bil_stride; ap.bil_off = bil_off; ap.out = out; ap.N = s->N; ap.H = s->H; ap.W = s->W; ap.mode = mode;
  { TraceScope ts_(ctx, "sr_apply_kernel"); sr_apply_kernel<<<cdiv((int64_t)s->N * s->H * s->W, 256), 256, 0, ctx->stream>>>(ap); }
  STC_CUDA(cudaGetLastError());
  ctx->launches++;
  return STC_OK;
}

int sr_forward_dev(stc_ctx* ctx, const float* x_dev, const float* bil_dev, int N, int H, int W, float* out_dev) {
  SrState* s = (SrState*)ctx->sr;
  if (!s || !s->ready) STC_FAIL(STC_ERR_STATE, "superresolve: weights not finalized");
  if (H < 3 || W < 3 || N < 1) STC_FAIL(STC_ERR_ARG, "superresolve: bad shape");
  int rc = sr_plan(ctx, s, N, H, W); if (rc) return rc;
  { TraceScope ts_(ctx, "sr_prep_kernel"); sr_prep_kernel<<<cdiv((int64_t)N * H * W, 256), 256, 0, ctx->stream>>>(x_dev, N, H, W, s->X.at(0), s->X.plane); }
  STC_CUDA(cudaGetLastError()); ctx->launches++;
  // Fused route (tcgen05 kernels): every convolution writes what the next one reads -- the fp16 activation with its reflect
  // border, the fp32 residual, and at the end tanh + bilinear -- from its accumulator.  The separate fp32 round trip
  // (sr_apply_kernel, below) moved 2.4x the bytes; it stays for the SIMT reference kernels (conv_impl 1) and as the A/B
  // reference (STC_SR_FUSE=0): both routes give identical bits (tests/test_gpu_preproc.py).
  const char* fuse_s = getenv("STC_SR_FUSE");                   // read per call: the A/B test flips it
  const bool fuse_env = !(fuse_s && atoi(fuse_s) == 0);
  if (fuse_env && ctx->conv_impl != 1) {
    if ((rc = sr_conv(ctx, s, 0, s->X, MODE_BIAS_RELU, &s->A, 1))) return rc;         // a (skip = a)
    if ((rc = sr_conv(ctx, s, 1, s->A, MODE_BIAS_RELU, &s->Bf, 0))) return rc;
    if ((rc = sr_conv(ctx, s, 2, s->Bf, MODE_BIAS, &s->A, 2))) return rc;             // b = a + 0.1*c
    if ((rc = sr_conv(ctx, s, 3, s->A, MODE_BIAS_RELU, &s->Bf, 0))) return rc;
    if ((rc = sr_conv(ctx, s, 4, s->Bf, MODE_BIAS, &s->A, 2))) return rc;             // c = b + 0.1*d
    // bil_dev == nullptr: the bilinear input is bands 4..9 of x itself (what superresolve_large_tile feeds, :104-105)
    if (bil_dev) return sr_conv(ctx, s, 5, s->A, MODE_BIAS, nullptr, 0, out_dev, bil_dev, 6, 0);
    return sr_conv(ctx, s, 5, s->A, MODE_BIAS, nullptr, 0, out_dev, x_dev, 10, 4);
  }
  if ((rc = sr_conv(ctx, s, 0, s->X, MODE_BIAS_RELU))) return rc;
  if ((rc = sr_apply(ctx, s, 0, 1, &s->A, nullptr, nullptr))) return rc;      // a (skip = a)
  if ((rc = sr_conv(ctx, s, 1, s->A, MODE_BIAS_RELU))) return rc;
  if ((rc = sr_apply(ctx, s, 0, 0, &s->Bf, nullptr, nullptr))) return rc;
  if ((rc = sr_conv(ctx, s, 2, s->Bf, MODE_BIAS))) return rc;
  if ((rc = sr_apply(ctx, s, 1, 0, &s->A, nullptr, nullptr))) return rc;      // b = a + 0.1*c
  if ((rc = sr_conv(ctx, s, 3, s->A, MODE_BIAS_RELU))) return rc;
  if ((rc = sr_apply(ctx, s, 0, 0, &s->Bf, nullptr, nullptr))) return rc;
  if ((rc = sr_conv(ctx, s, 4, s->Bf, MODE_BIAS))) return rc;
  if ((rc = sr_apply(ctx, s, 1, 0, &s->A, nullptr, nullptr))) return rc;      // c = b + 0.1*d
  if ((rc = sr_conv(ctx, s, 5, s->A, MODE_BIAS))) return rc;
  // bil_dev == nullptr: the bilinear input is bands 4..9 of x itself (what superresolve_large_tile feeds, :104-105)
  if (bil_dev) rc = sr_apply(ctx, s, 2, 0, nullptr, bil_dev, out_dev);
  else rc = sr_apply(ctx, s, 2, 0, nullptr, x_dev, out_dev, 10, 4);
  if (rc) return rc;
  return STC_OK;
}

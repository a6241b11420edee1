// HBM-bound NHWC fp16 helpers of the YOLOv9-E graph (ref:util/yolov9.py:120-121 runs them inside the
// TorchScript archive): ADown pooling, SPPELAN max-pool, nearest upsample, CBFuse multi-scale sum.
// Every tensor is a channel slice (pointer + pixel stride `ld`) so concats are never materialised.
// 8 channels (one 16-byte vector) per thread, channels fastest => fully coalesced.
#include "b2p_internal.h"
#include <cuda_fp16.h>

namespace b2p {

struct H8 { uint4 v; };

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __half22float2(h[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __half2* h = reinterpret_cast<__half2*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return v;
}
__device__ __forceinline__ uint4 ld8(const __half* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void st8(__half* p, const uint4& v) { *reinterpret_cast<uint4*>(p) = v; }
// fp16x3 ("parity-grade") maps keep every value as an fp16 hi/lo pair: hi at p, lo at p + lo (lo = channel count of the
// owning buffer; 0 = plain fp16 map).  hi + lo is exact in fp32 (22 significant bits).
__device__ __forceinline__ void ldf8(const __half* p, long long lo, float (&f)[8]) {
  unpack8(ld8(p), f);
  if (lo) {
    float g[8];
    unpack8(ld8(p + lo), g);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] += g[k];
  }
}
__device__ __forceinline__ void stf8(__half* p, long long lo, const float (&f)[8]) {
  const uint4 hi = pack8(f);
  st8(p, hi);
  if (lo) {
    float h[8], r[8];
    unpack8(hi, h);
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = f[k] - h[k];
    st8(p + lo, pack8(r));
  }
}

// ADown front half (common.py::ADown.forward): a = avg_pool2d(x, 2, 1) (size (H-1)x(W-1)); x1 = a[:, :C/2],
// x2 = max_pool2d(a[:, C/2:], 3, 2, 1).  x1 is stored as an HxW map whose last row/column are zero, which is
// exactly the zero padding the following 3x3 stride-2 pad-1 conv would see (keeps H, W even for the TMA view).
__global__ void adown_pool_kernel(const __half* __restrict__ x, long long ldx, int B, int H, int W, int C,
                                  __half* __restrict__ x1, long long ld1, __half* __restrict__ x2, long long ld2,
                                  long long lox, long long lo1, long long lo2) {
  pdl_wait();   // PDL: inputs come from the previous kernel in the stream
  const int c8n = C / 16;   // vectors per half
  const long long n1 = (long long)B * H * W * c8n;
  const int Ho = H / 2, Wo = W / 2;
  const long long n2 = (long long)B * Ho * Wo * c8n;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n1 + n2; i += (long long)gridDim.x * blockDim.x) {
    if (i < n1) {
      const int cv = int(i % c8n);
      long long p = i / c8n;
      const int xx = int(p % W); p /= W;
      const int yy = int(p % H);
      const int b = int(p / H);
      float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (yy < H - 1 && xx < W - 1) {
        const __half* s = x + (((long long)b * H + yy) * W + xx) * ldx + cv * 8;
        float a[8], bq[8], c[8], d[8];
        ldf8(s, lox, a); ldf8(s + ldx, lox, bq); ldf8(s + (long long)W * ldx, lox, c); ldf8(s + (long long)(W + 1) * ldx, lox, d);
#pragma unroll
        for (int k = 0; k < 8; ++k) o[k] = ((a[k] + bq[k]) + (c[k] + d[k])) * 0.25f;
      }
      stf8(x1 + (((long long)b * H + yy) * W + xx) * ld1 + cv * 8, lo1, o);
    } else {
      const long long j = i - n1;
      const int cv = int(j % c8n);
      long long p = j / c8n;
      const int ox = int(p % Wo); p /= Wo;
      const int oy = int(p % Ho);
      const int b = int(p / Ho);
      float m[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
      for (int dy = -1; dy <= 1; ++dy) {
        const int yy = 2 * oy + dy;
        if (yy < 0 || yy >= H - 1) continue;
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = 2 * ox + dx;
          if (xx < 0 || xx >= W - 1) continue;
          const __half* s = x + (((long long)b * H + yy) * W + xx) * ldx + C / 2 + cv * 8;
          float a[8], bq[8], c[8], d[8];
          ldf8(s, lox, a); ldf8(s + ldx, lox, bq); ldf8(s + (long long)W * ldx, lox, c); ldf8(s + (long long)(W + 1) * ldx, lox, d);
#pragma unroll
          for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], ((a[k] + bq[k]) + (c[k] + d[k])) * 0.25f);
        }
      }
      stf8(x2 + (((long long)b * Ho + oy) * Wo + ox) * ld2 + cv * 8, lo2, m);
    }
  }
}

// Plain fp16 maps (the fast detector mode): the two halves of adown_pool_kernel as separate launches, one item per thread,
// 32-bit index arithmetic.  The x2 item of the combined kernel walks nine 2x2 windows one after the other (nine dependent
// L2 round trips, and its register count sets the occupancy of the x1 items); here its 4x4 input patch (16 vector loads) is
// requested at once.  hs = horizontal pair sums of a patch row; an average is (hs[row] + hs[row + 1]) * 0.25 =
// ((a + b) + (c + d)) * 0.25: the same association as the combined kernel => identical results.
__global__ void __launch_bounds__(256) adown_avg_kernel(const __half* __restrict__ x, long long ldx, int B, int H, int W, int C,
                                                        __half* __restrict__ x1, long long ld1) {
  pdl_wait();
  const unsigned c8n = unsigned(C) / 16u;
  const unsigned n1 = unsigned(B) * H * W * c8n;
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n1) return;
  const unsigned cv = i % c8n;
  unsigned p = i / c8n;
  const unsigned xx = p % unsigned(W); p /= unsigned(W);
  const unsigned yy = p % unsigned(H);
  const unsigned b = p / unsigned(H);
  float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (yy < unsigned(H - 1) && xx < unsigned(W - 1)) {
    const __half* s = x + (((long long)b * H + yy) * W + xx) * ldx + cv * 8;
    const uint4 ra = ld8(s), rb = ld8(s + ldx), rc = ld8(s + (long long)W * ldx), rd = ld8(s + (long long)(W + 1) * ldx);
    float a[8], bq[8], c[8], d[8];
    unpack8(ra, a); unpack8(rb, bq); unpack8(rc, c); unpack8(rd, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = ((a[k] + bq[k]) + (c[k] + d[k])) * 0.25f;
  }
  st8(x1 + (((long long)b * H + yy) * W + xx) * ld1 + cv * 8, pack8(o));
}

__global__ void __launch_bounds__(256) adown_max_kernel(const __half* __restrict__ x, long long ldx, int B, int H, int W, int C,
                                                        __half* __restrict__ x2, long long ld2) {
  pdl_wait();
  const unsigned c8n = unsigned(C) / 16u;
  const int Ho = H / 2, Wo = W / 2;
  const unsigned n2 = unsigned(B) * Ho * Wo * c8n;
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n2) return;
  const unsigned cv = i % c8n;
  unsigned p = i / c8n;
  const int ox = int(p % unsigned(Wo)); p /= unsigned(Wo);
  const int oy = int(p % unsigned(Ho));
  const int b = int(p / unsigned(Ho));
  uint4 raw[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int yy = 2 * oy - 1 + r;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int xx = 2 * ox - 1 + c;
      raw[r][c] = (yy >= 0 && yy < H && xx >= 0 && xx < W)
                      ? ld8(x + (((long long)b * H + yy) * W + xx) * ldx + C / 2 + cv * 8) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  float m[8], hp[3][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float pr[4][8], hs[3][8];
#pragma unroll
    for (int c = 0; c < 4; ++c) unpack8(raw[r][c], pr[c]);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < 8; ++k) hs[c][k] = pr[c][k] + pr[c + 1][k];
    }
    if (r > 0) {
      const int ay = 2 * oy - 2 + r;               // average row: inputs ay, ay + 1
      if (ay >= 0 && ay < H - 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int ax = 2 * ox - 1 + c;
          if (ax >= 0 && ax < W - 1) {
#pragma unroll
            for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], (hp[c][k] + hs[c][k]) * 0.25f);
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int k = 0; k < 8; ++k) hp[c][k] = hs[c][k];
    }
  }
  st8(x2 + (((long long)b * Ho + oy) * Wo + ox) * ld2 + cv * 8, pack8(m));
}

// MaxPool2d(k, stride 1, pad k/2) on a channel slice (SPPELAN, k = 5).
__global__ void maxpool_s1_kernel(const __half* __restrict__ x, long long ldx, int B, int H, int W, int C, int k,
                                  __half* __restrict__ y, long long ldy, long long lox, long long loy) {
  pdl_wait();   // PDL: inputs come from the previous kernel in the stream
  const int cvn = C / 8;
  const long long n = (long long)B * H * W * cvn;
  const int r = k / 2;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cv = int(i % cvn);
    long long p = i / cvn;
    const int xx = int(p % W); p /= W;
    const int yy = int(p % H);
    const int b = int(p / H);
    if (lox) {   // hi/lo pairs: max over the exact fp32 values
      float mf[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) mf[q] = -INFINITY;
      for (int dy = -r; dy <= r; ++dy) {
        const int sy = yy + dy;
        if (sy < 0 || sy >= H) continue;
        for (int dx = -r; dx <= r; ++dx) {
          const int sx = xx + dx;
          if (sx < 0 || sx >= W) continue;
          float t[8];
          ldf8(x + (((long long)b * H + sy) * W + sx) * ldx + cv * 8, lox, t);
#pragma unroll
          for (int q = 0; q < 8; ++q) mf[q] = fmaxf(mf[q], t[q]);
        }
      }
      stf8(y + (((long long)b * H + yy) * W + xx) * ldy + cv * 8, loy, mf);
      continue;
    }
    __half2 m[4];
    const __half2 ninf = __float2half2_rn(-INFINITY);
#pragma unroll
    for (int q = 0; q < 4; ++q) m[q] = ninf;
    for (int dy = -r; dy <= r; ++dy) {
      const int sy = yy + dy;
      if (sy < 0 || sy >= H) continue;
      for (int dx = -r; dx <= r; ++dx) {
        const int sx = xx + dx;
        if (sx < 0 || sx >= W) continue;
        const uint4 v = ld8(x + (((long long)b * H + sy) * W + sx) * ldx + cv * 8);
        const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int q = 0; q < 4; ++q) m[q] = __hmax2(m[q], h[q]);
      }
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int q = 0; q < 4; ++q) oh[q] = m[q];
    st8(y + (((long long)b * H + yy) * W + xx) * ldy + cv * 8, o);
  }
}

// nn.Upsample(scale_factor=2, mode='nearest') into a channel slice.
__global__ void upsample2x_kernel(const __half* __restrict__ x, long long ldx, int B, int H, int W, int C,
                                  __half* __restrict__ y, long long ldy, long long lox, long long loy) {
  pdl_wait();   // PDL: inputs come from the previous kernel in the stream
  const int cvn = C / 8;
  const int Ho = 2 * H, Wo = 2 * W;
  const long long n = (long long)B * Ho * Wo * cvn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cv = int(i % cvn);
    long long p = i / cvn;
    const int xx = int(p % Wo); p /= Wo;
    const int yy = int(p % Ho);
    const int b = int(p / Ho);
    st8(y + (((long long)b * Ho + yy) * Wo + xx) * ldy + cv * 8,
        ld8(x + (((long long)b * H + (yy >> 1)) * W + (xx >> 1)) * ldx + cv * 8));
    if (lox)
      st8(y + (((long long)b * Ho + yy) * Wo + xx) * ldy + loy + cv * 8,
          ld8(x + (((long long)b * H + (yy >> 1)) * W + (xx >> 1)) * ldx + lox + cv * 8));
  }
}

struct FuseSrc { const __half* p; long long ld; int shift; int H, W; long long lo; };
struct FuseArgs { FuseSrc s[5]; int n; };

// CBFuse (common.py::CBFuse): out = sum_i nearest_upsample(src_i) + last, summed in that order in fp32.
__global__ void cbfuse_kernel(FuseArgs a, const __half* __restrict__ last, long long ldl, int B, int H, int W, int C,
                              __half* __restrict__ y, long long ldy, long long lol, long long loy) {
  pdl_wait();   // PDL: inputs come from the previous kernel in the stream
  const int cvn = C / 8;
  const long long n = (long long)B * H * W * cvn;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int cv = int(i % cvn);
    long long p = i / cvn;
    const int xx = int(p % W); p /= W;
    const int yy = int(p % H);
    const int b = int(p / H);
    float acc[8], t[8];
    bool first = true;
    for (int s = 0; s < a.n; ++s) {
      const FuseSrc& f = a.s[s];
      ldf8(f.p + (((long long)b * f.H + (yy >> f.shift)) * f.W + (xx >> f.shift)) * f.ld + cv * 8, f.lo, t);
      if (first) {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = t[k];
        first = false;
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += t[k];
      }
    }
    ldf8(last + (((long long)b * H + yy) * W + xx) * ldl + cv * 8, lol, t);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = first ? t[k] : acc[k] + t[k];
    stf8(y + (((long long)b * H + yy) * W + xx) * ldy + cv * 8, loy, acc);
  }
}

// Plain fp16 maps (the fast detector mode), NS sources known at compile time: ALL NS + 1 vector loads of an item are issued
// before the first use.  The loop above waits for each source before it requests the next (ncu: long-scoreboard stalls,
// 2.3 TB/s); one item per thread, 32-bit index arithmetic.  Same fp32 sums in the same order.
template <int NS>
__global__ void __launch_bounds__(256) cbfuse_batched_kernel(FuseArgs a, const __half* __restrict__ last, long long ldl, int B, int H, int W,
                                                             int C, __half* __restrict__ y, long long ldy) {
  pdl_wait();
  const unsigned cvn = unsigned(C) / 8u;
  const unsigned n = unsigned(B) * H * W * cvn;
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  if (i >= n) return;
  const unsigned cv = i % cvn;
  unsigned p = i / cvn;
  const unsigned xx = p % unsigned(W); p /= unsigned(W);
  const unsigned yy = p % unsigned(H);
  const unsigned b = p / unsigned(H);
  uint4 raw[NS + 1];
#pragma unroll
  for (int s = 0; s < NS; ++s) {
    const FuseSrc& f = a.s[s];
    raw[s] = ld8(f.p + (((long long)b * f.H + (yy >> f.shift)) * f.W + (xx >> f.shift)) * f.ld + cv * 8);
  }
  raw[NS] = ld8(last + (((long long)b * H + yy) * W + xx) * ldl + cv * 8);
  float acc[8], t[8];
  unpack8(raw[0], acc);
#pragma unroll
  for (int s = 1; s <= NS; ++s) {
    unpack8(raw[s], t);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += t[k];
  }
  st8(y + (((long long)b * H + yy) * W + xx) * ldy + cv * 8, pack8(acc));
}

static inline int grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = 148LL * 16;
  return int(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace b2p

using namespace b2p;

extern "C" {

// `_x3` forms: every map is an fp16 hi/lo pair (lo plane `lo*` elements after the hi plane); the arithmetic is done
// on the exact fp32 sums and the result is re-split.  lo == 0 everywhere reproduces the plain fp16 kernels.
int b2p_adown_pool_x3(const void* x, long long ldx, int B, int H, int W, int C, void* x1, long long ld1, void* x2,
                      long long ld2, long long lox, long long lo1, long long lo2, cudaStream_t st) {
  if ((H & 1) || (W & 1) || (C % 16)) return set_error("adown_pool: H, W must be even and C a multiple of 16");
  if ((lox != 0) != (lo1 != 0) || (lox != 0) != (lo2 != 0) || (lox | lo1 | lo2) % 8) return set_error("adown_pool: lo planes must be all set (8-aligned) or all 0");
  const long long n = (long long)B * H * W * (C / 16) + (long long)B * (H / 2) * (W / 2) * (C / 16);
  if (lox == 0 && n < (1LL << 31) - 256) {
    const long long n1 = (long long)B * H * W * (C / 16), n2 = n - n1;
    launch_pdl(adown_avg_kernel, dim3(unsigned((n1 + 255) / 256)), dim3(256), 0, st, (const __half*)x, ldx, B, H, W, C, (__half*)x1, ld1);
    B2P_CHECK_LAUNCH();
    launch_pdl(adown_max_kernel, dim3(unsigned((n2 + 255) / 256)), dim3(256), 0, st, (const __half*)x, ldx, B, H, W, C, (__half*)x2, ld2);
    B2P_CHECK_LAUNCH();
    return 0;
  }
  launch_pdl(adown_pool_kernel, dim3(grid_for(n, 256)), dim3(256), 0, st, (const __half*)x, ldx, B, H, W, C, (__half*)x1, ld1, (__half*)x2, ld2,
             lox, lo1, lo2);
  B2P_CHECK_LAUNCH();
  return 0;
}
int b2p_adown_pool(const void* x, long long ldx, int B, int H, int W, int C, void* x1, long long ld1, void* x2,
                   long long ld2, cudaStream_t st) {
  return b2p_adown_pool_x3(x, ldx, B, H, W, C, x1, ld1, x2, ld2, 0, 0, 0, st);
}

int b2p_maxpool_s1_x3(const void* x, long long ldx, int B, int H, int W, int C, int k, void* y, long long ldy, long long lox,
                      long long loy, cudaStream_t st) {
  if (C % 8) return set_error("maxpool_s1: C must be a multiple of 8");
  if ((lox != 0) != (loy != 0) || (lox | loy) % 8) return set_error("maxpool_s1: lo planes must be both set (8-aligned) or both 0");
  const long long n = (long long)B * H * W * (C / 8);
  launch_pdl(maxpool_s1_kernel, dim3(grid_for(n, 256)), dim3(256), 0, st, (const __half*)x, ldx, B, H, W, C, k, (__half*)y, ldy, lox, loy);
  B2P_CHECK_LAUNCH();
  return 0;
}
int b2p_maxpool_s1(const void* x, long long ldx, int B, int H, int W, int C, int k, void* y, long long ldy, cudaStream_t st) {
  return b2p_maxpool_s1_x3(x, ldx, B, H, W, C, k, y, ldy, 0, 0, st);
}

int b2p_upsample2x_x3(const void* x, long long ldx, int B, int H, int W, int C, void* y, long long ldy, long long lox,
                      long long loy, cudaStream_t st) {
  if (C % 8) return set_error("upsample2x: C must be a multiple of 8");
  if ((lox != 0) != (loy != 0) || (lox | loy) % 8) return set_error("upsample2x: lo planes must be both set (8-aligned) or both 0");
  const long long n = (long long)B * 4 * H * W * (C / 8);
  launch_pdl(upsample2x_kernel, dim3(grid_for(n, 256)), dim3(256), 0, st, (const __half*)x, ldx, B, H, W, C, (__half*)y, ldy, lox, loy);
  B2P_CHECK_LAUNCH();
  return 0;
}
int b2p_upsample2x(const void* x, long long ldx, int B, int H, int W, int C, void* y, long long ldy, cudaStream_t st) {
  return b2p_upsample2x_x3(x, ldx, B, H, W, C, y, ldy, 0, 0, st);
}

// Explicit im2col of an NHWC fp16 map for a 3x3 / pad 1 conv: out[(b,yo,xo)][tap*C + c] = x[b][yo*s+ky-1][xo*s+kx-1][c] (0 outside).
// Used where the implicit-GEMM tile would be mostly padding (DaViT stage 2/3 patch-embed convs on 4x4 / 2x2 output maps of the
// 64x64-crop mode: a 128-pixel tile holds 16 / 4 real pixels); the result feeds the plain GEMM.  HBM-bound row copies (uint4).
// halves == 2: fp16x3 operands -- pixels are [hi(C) | lo(C)] and the output row is [hi: 9*C | lo: 9*C].
__global__ void im2col3x3_kernel(const __half* __restrict__ x, long long ldx, int B, int H, int W, int C, int s, int Ho, int Wo,
                                 int halves, __half* __restrict__ out) {
  pdl_wait();
  const int c8 = C >> 3;
  const long long total = (long long)B * Ho * Wo * halves * 9 * c8;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = int(i % c8);
    long long r = i / c8;
    const int tap = int(r % 9); r /= 9;
    const int half = int(r % halves); r /= halves;
    const int xo = int(r % Wo); r /= Wo;
    const int yo = int(r % Ho);
    const int b = int(r / Ho);
    const int y = yo * s + tap / 3 - 1, xx = xo * s + tap % 3 - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (y >= 0 && y < H && xx >= 0 && xx < W)
      v = *reinterpret_cast<const uint4*>(x + (((long long)b * H + y) * W + xx) * ldx + half * C + c * 8);
    reinterpret_cast<uint4*>(out)[i] = v;
  }
}

int b2p_im2col3x3(const void* x, long long ldx, int B, int H, int W, int C, int stride, int halves, void* out, cudaStream_t st) {
  if (C % 8 || ldx % 8) return set_error("im2col3x3: C and the pixel stride must be multiples of 8");
  if (stride != 1 && stride != 2) return set_error("im2col3x3: stride 1 or 2");
  if (halves != 1 && halves != 2) return set_error("im2col3x3: halves is 1 (plain) or 2 (fp16x3 [hi | lo] pixels)");
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  const long long n = (long long)B * Ho * Wo * halves * 9 * (C / 8);
  if (n == 0) return 0;
  launch_pdl(im2col3x3_kernel, dim3(grid_for(n, 256)), dim3(256), 0, st, (const __half*)x, ldx, B, H, W, C, stride, Ho, Wo, halves,
             (__half*)out);
  B2P_CHECK_LAUNCH();
  return 0;
}

// srcs[i]: pointer to the selected CBLinear split (already offset to its first channel), lds[i] its pixel stride,
// shifts[i] = log2(target size / source size).
int b2p_cbfuse_x3(int nsrc, const void* const* srcs, const long long* lds, const int* shifts, const long long* los, const void* last,
                  long long ldl, int B, int H, int W, int C, void* y, long long ldy, long long lol, long long loy, cudaStream_t st) {
  if (nsrc < 0 || nsrc > 5) return set_error("cbfuse: at most 5 upsampled sources");
  if (C % 8) return set_error("cbfuse: C must be a multiple of 8");
  if ((lol != 0) != (loy != 0) || (lol | loy) % 8) return set_error("cbfuse: lo planes must be all set (8-aligned) or all 0");
  FuseArgs a{};
  a.n = nsrc;
  for (int i = 0; i < nsrc; ++i) {
    a.s[i].p = (const __half*)srcs[i];
    a.s[i].ld = lds[i];
    a.s[i].shift = shifts[i];
    a.s[i].H = H >> shifts[i];
    a.s[i].W = W >> shifts[i];
    a.s[i].lo = los ? los[i] : 0;
    if ((a.s[i].lo != 0) != (lol != 0) || a.s[i].lo % 8) return set_error("cbfuse: lo planes must be all set (8-aligned) or all 0");
  }
  const long long n = (long long)B * H * W * (C / 8);
  if (lol == 0 && nsrc >= 1 && n < (1LL << 31) - 256) {
    const dim3 grid(unsigned((n + 255) / 256)), blk(256);
    const __half* lp = (const __half*)last;
    __half* yp = (__half*)y;
    switch (nsrc) {
      case 1: launch_pdl(cbfuse_batched_kernel<1>, grid, blk, 0, st, a, lp, ldl, B, H, W, C, yp, ldy); break;
      case 2: launch_pdl(cbfuse_batched_kernel<2>, grid, blk, 0, st, a, lp, ldl, B, H, W, C, yp, ldy); break;
      case 3: launch_pdl(cbfuse_batched_kernel<3>, grid, blk, 0, st, a, lp, ldl, B, H, W, C, yp, ldy); break;
      case 4: launch_pdl(cbfuse_batched_kernel<4>, grid, blk, 0, st, a, lp, ldl, B, H, W, C, yp, ldy); break;
      default: launch_pdl(cbfuse_batched_kernel<5>, grid, blk, 0, st, a, lp, ldl, B, H, W, C, yp, ldy); break;
    }
    B2P_CHECK_LAUNCH();
    return 0;
  }
  launch_pdl(cbfuse_kernel, dim3(grid_for(n, 256)), dim3(256), 0, st, a, (const __half*)last, ldl, B, H, W, C, (__half*)y, ldy, lol, loy);
  B2P_CHECK_LAUNCH();
  return 0;
}
int b2p_cbfuse(int nsrc, const void* const* srcs, const long long* lds, const int* shifts, const void* last,
               long long ldl, int B, int H, int W, int C, void* y, long long ldy, cudaStream_t st) {
  return b2p_cbfuse_x3(nsrc, srcs, lds, shifts, nullptr, last, ldl, B, H, W, C, y, ldy, 0, 0, st);
}

}  // extern "C"

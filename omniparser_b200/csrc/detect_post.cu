// Detector post-processing on device, replacing the eager torch ops + host syncs of
// ref:util/yolov9.py:89-108 (_decode), :123-129 (class max, conf filter, un-letterbox) and :131-136
// (torchvision batched_nms, [:max_det], clamp).  Integer results (candidate order, kept indices) are
// bit-exact with the reference CPU path given the same head tensors; the fp32 arithmetic uses explicit
// round-to-nearest intrinsics (no FMA contraction) in the reference's operation order.
//
// Candidate order = ascending anchor index (scale-major 8 -> 16 -> 32, row-major y then x), the order
// boolean-mask indexing produces at ref:util/yolov9.py:124-127.  NMS order = descending score, ties by
// ascending candidate index (torchvision sorts with stable=True), suppress when IoU > thr (strict).
#include "b2p_internal.h"
#include <atomic>
#include <math.h>

namespace b2p {

struct DecodeArgs {
  const float* cls[3];   // [B][H*W][nc] class logits per scale
  const float* box[3];   // [B][H*W][64] DFL logits per scale, channel = side*16 + bin
  int H[3], W[3];
  int nc, B, cap;
  float conf;
  const float* pad_l;    // [B] letterbox pad_left, pad_top, scale (as float32, ref:util/yolov9.py:76-80)
  const float* pad_t;
  const float* scale;
  float* cand_box;       // [B][cap][4]
  float* cand_score;     // [B][cap]
  int* cand_cls;         // [B][cap]
  int* cand_count;       // [B]  (total candidates, may exceed cap -> caller must check)
  float* dense_ltrb;     // optional [B][A][4] (stride units) for parity checks
  float* dense_score;    // optional [B][A][nc] sigmoid scores
};

__global__ void __launch_bounds__(1024) yolo_decode_kernel(DecodeArgs a) {
  const int b = blockIdx.x;
  __shared__ int warp_cnt[32];
  __shared__ int base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) base_s = 0;
  __syncthreads();
  int A = 0;
  for (int s = 0; s < 3; ++s) A += a.H[s] * a.W[s];
  const float padl = a.pad_l[b], padt = a.pad_t[b], sc = a.scale[b];
  for (int a0 = 0; a0 < A; a0 += blockDim.x) {
    const int ai = a0 + threadIdx.x;
    bool keep = false;
    float bx[4] = {0, 0, 0, 0}, best = 0.f;
    int bc = 0;
    if (ai < A) {
      int s = 0, off = 0;
      while (ai >= off + a.H[s] * a.W[s]) { off += a.H[s] * a.W[s]; ++s; }
      const int local = ai - off;
      const int gy = local / a.W[s], gx = local - gy * a.W[s];
      const float stride = float(8 << s);
      const long long pix = (long long)b * a.H[s] * a.W[s] + local;
      // class scores: sigmoid, max over classes (first maximum wins)
      const float* cl = a.cls[s] + pix * a.nc;
      best = -1.f;
      for (int c = 0; c < a.nc; ++c) {
        const float p = 1.0f / (1.0f + expf(-cl[c]));
        if (a.dense_score) a.dense_score[((long long)b * A + ai) * a.nc + c] = p;
        if (p > best) { best = p; bc = c; }
      }
      // DFL: softmax over 16 bins, expectation (stride units)
      const float* bl = a.box[s] + pix * 64;
      float d[4];
#pragma unroll
      for (int side = 0; side < 4; ++side) {
        float v[16], m = -INFINITY;
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i] = bl[side * 16 + i]; m = fmaxf(m, v[i]); }
        float sum = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) { v[i] = expf(v[i] - m); sum += v[i]; }
        float e = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) e += (v[i] / sum) * float(i);
        d[side] = e;
        if (a.dense_ltrb) a.dense_ltrb[((long long)b * A + ai) * 4 + side] = e;
      }
      // ref:util/yolov9.py:97-106: dist*stride; anchors (grid+0.5)*stride; [a - lt, a + rb]
      const float ax = __fmul_rn(float(gx) + 0.5f, stride), ay = __fmul_rn(float(gy) + 0.5f, stride);
      const float x1 = __fsub_rn(ax, __fmul_rn(d[0], stride)), y1 = __fsub_rn(ay, __fmul_rn(d[1], stride));
      const float x2 = __fadd_rn(ax, __fmul_rn(d[2], stride)), y2 = __fadd_rn(ay, __fmul_rn(d[3], stride));
      keep = best > a.conf;   // strict, ref:util/yolov9.py:124
      // ref:util/yolov9.py:128-129: (x - pad) / scale
      bx[0] = __fdiv_rn(__fsub_rn(x1, padl), sc);
      bx[1] = __fdiv_rn(__fsub_rn(y1, padt), sc);
      bx[2] = __fdiv_rn(__fsub_rn(x2, padl), sc);
      bx[3] = __fdiv_rn(__fsub_rn(y2, padt), sc);
    }
    // order-preserving compaction
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    int before = 0, total = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      const int c = warp_cnt[w];
      if (w < warp) before += c;
      total += c;
    }
    const int pos = base_s + before + __popc(m & ((1u << lane) - 1));
    if (keep && pos < a.cap) {
      float* cb = a.cand_box + ((long long)b * a.cap + pos) * 4;
      cb[0] = bx[0]; cb[1] = bx[1]; cb[2] = bx[2]; cb[3] = bx[3];
      a.cand_score[(long long)b * a.cap + pos] = best;
      a.cand_cls[(long long)b * a.cap + pos] = bc;
    }
    __syncthreads();
    if (threadIdx.x == 0) base_s += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) a.cand_count[b] = base_s;
}

// ------------------------------------------------------------------------------------------- NMS
static constexpr int kNmsThreads = 512;
static constexpr int kNmsMaxKeep = 1024;

struct NmsArgs {
  const float* box;      // [B][cap][4]
  const float* score;    // [B][cap]
  const int* cls;        // [B][cap]
  const int* count;      // [B]
  int cap, max_det, sort_cap;
  float thr;             // largest float <= the (double) IoU threshold, see host
  const float* img_w;    // [B] clamp bounds (ref:util/yolov9.py:134-135)
  const float* img_h;
  int* keep_idx;         // [B][max_det]  candidate indices in NMS order
  float* out_box;        // [B][max_det][4] clamped
  float* out_score;      // [B][max_det]
  int* out_count;        // [B]
};

__device__ __forceinline__ bool iou_gt(const float4& p, float pa, const float4& q, float qa, float thr) {
  // torchvision nms: w = max(0, xx2 - xx1); inter = w*h; ovr = inter / (iarea + areas[j] - inter); ovr > thr
  const float xx1 = fmaxf(p.x, q.x), yy1 = fmaxf(p.y, q.y), xx2 = fminf(p.z, q.z), yy2 = fminf(p.w, q.w);
  const float w = fmaxf(0.f, __fsub_rn(xx2, xx1)), h = fmaxf(0.f, __fsub_rn(yy2, yy1));
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(pa, qa), inter));
  return ovr > thr;
}

__global__ void __launch_bounds__(kNmsThreads) batched_nms_kernel(NmsArgs a) {
  extern __shared__ unsigned long long keys[];          // [sort_cap], then the carved arrays below
  unsigned char* sp = reinterpret_cast<unsigned char*>(keys + a.sort_cap);
  float4* kbox = reinterpret_cast<float4*>(sp);            sp += sizeof(float4) * kNmsMaxKeep;
  float4* cbox = reinterpret_cast<float4*>(sp);            sp += sizeof(float4) * kNmsThreads;
  float* karea = reinterpret_cast<float*>(sp);             sp += sizeof(float) * kNmsMaxKeep;
  int* kcls = reinterpret_cast<int*>(sp);                  sp += sizeof(int) * kNmsMaxKeep;
  int* kidx = reinterpret_cast<int*>(sp);                  sp += sizeof(int) * kNmsMaxKeep;
  float* carea = reinterpret_cast<float*>(sp);             sp += sizeof(float) * kNmsThreads;
  int* ccls = reinterpret_cast<int*>(sp);                  sp += sizeof(int) * kNmsThreads;
  unsigned (*cmask)[kNmsThreads / 32] = reinterpret_cast<unsigned (*)[kNmsThreads / 32]>(sp);
  __shared__ unsigned cpre[kNmsThreads / 32];
  __shared__ int nkept_s;
  __shared__ float red[kNmsThreads / 32];
  __shared__ float maxc_s;

  const int b = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  int n = a.count[b];
  if (n > a.cap) n = a.cap;
  if (n > a.sort_cap) n = a.sort_cap;   // host rejects this case before launch when it can; defensive
  const float* box = a.box + (long long)b * a.cap * 4;
  const float* score = a.score + (long long)b * a.cap;
  const int* cls = a.cls + (long long)b * a.cap;

  int P = 1;
  while (P < n) P <<= 1;
  for (int i = t; i < P; i += blockDim.x) {
    unsigned long long k = ~0ull;
    if (i < n) {
      unsigned u = __float_as_uint(score[i]);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);   // monotone map float -> uint
      k = ((unsigned long long)(~u) << 32) | (unsigned)i;
    }
    keys[i] = k;
  }
  // coordinate trick (tv:ops/boxes.py _batched_nms_coordinate_trick) when numel <= 4000, else per-class
  const bool trick = (n * 4 <= 4000);
  float mx = -INFINITY;
  if (trick)
    for (int i = t; i < n * 4; i += blockDim.x) mx = fmaxf(mx, box[i]);
  for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  if (t == 0) nkept_s = 0;
  __syncthreads();
  if (t == 0) {
    float m = red[0];
    for (int w = 1; w < kNmsThreads / 32; ++w) m = fmaxf(m, red[w]);
    maxc_s = __fadd_rn(m, 1.0f);
  }
  // bitonic sort ascending (descending score, ascending index)
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = t; i < P; i += blockDim.x) {
        const int ixj = i ^ j;
        if (ixj > i) {
          const unsigned long long x = keys[i], y = keys[ixj];
          const bool up = ((i & k) == 0);
          if ((x > y) == up) { keys[i] = y; keys[ixj] = x; }
        }
      }
    }
  }
  __syncthreads();
  const float offs_unit = maxc_s;

  for (int c0 = 0; c0 < n; c0 += kNmsThreads) {
    const int cn = min(kNmsThreads, n - c0);
    const int nk = nkept_s;
    bool dead = true;
    float4 bq = make_float4(0, 0, 0, 0);
    float ar = 0.f;
    int mycls = 0;
    if (t < cn) {
      const int j = int(keys[c0 + t] & 0xffffffffu);
      mycls = cls[j];
      const float off = trick ? __fmul_rn(float(mycls), offs_unit) : 0.f;
      bq = make_float4(__fadd_rn(box[4 * j], off), __fadd_rn(box[4 * j + 1], off), __fadd_rn(box[4 * j + 2], off),
                       __fadd_rn(box[4 * j + 3], off));
      ar = __fmul_rn(__fsub_rn(bq.z, bq.x), __fsub_rn(bq.w, bq.y));
      dead = false;
      for (int k = 0; k < nk; ++k) {
        if ((trick || kcls[k] == mycls) && iou_gt(kbox[k], karea[k], bq, ar, a.thr)) { dead = true; break; }
      }
    }
    cbox[t] = bq; carea[t] = ar; ccls[t] = mycls;
    const unsigned dm = __ballot_sync(0xffffffffu, dead);
    if (lane == 0) cpre[warp] = dm;
    __syncthreads();
    // pairwise mask inside the chunk: bit u of row t set if t suppresses u (u > t)
    for (int w = 0; w < kNmsThreads / 32; ++w) {
      unsigned bits = 0;
      if (t < cn && !dead && (w * 32 + 31) > t) {
        for (int q = 0; q < 32; ++q) {
          const int u = w * 32 + q;
          if (u > t && u < cn && (trick || ccls[u] == mycls) && iou_gt(bq, ar, cbox[u], carea[u], a.thr)) bits |= 1u << q;
        }
      }
      cmask[t][w] = bits;
    }
    __syncthreads();
    if (warp == 0) {
      unsigned removed = (lane < kNmsThreads / 32) ? cpre[lane] : 0u;
      int nkl = nk;
      for (int i = 0; i < cn && nkl < a.max_det; ++i) {
        const unsigned wbits = __shfl_sync(0xffffffffu, removed, i >> 5);
        if (!((wbits >> (i & 31)) & 1u)) {
          if (lane == 0) {
            kbox[nkl] = cbox[i]; karea[nkl] = carea[i]; kcls[nkl] = ccls[i];
            kidx[nkl] = int(keys[c0 + i] & 0xffffffffu);
          }
          ++nkl;
          if (lane < kNmsThreads / 32) removed |= cmask[i][lane];
        }
      }
      if (lane == 0) nkept_s = nkl;
    }
    __syncthreads();
    if (nkept_s >= a.max_det) break;
  }
  const int nk = nkept_s;
  const float W = a.img_w[b], H = a.img_h[b];
  for (int k = t; k < nk; k += blockDim.x) {
    const int j = kidx[k];
    a.keep_idx[(long long)b * a.max_det + k] = j;
    float* o = a.out_box + ((long long)b * a.max_det + k) * 4;
    o[0] = fminf(fmaxf(box[4 * j], 0.f), W);
    o[1] = fminf(fmaxf(box[4 * j + 1], 0.f), H);
    o[2] = fminf(fmaxf(box[4 * j + 2], 0.f), W);
    o[3] = fminf(fmaxf(box[4 * j + 3], 0.f), H);
    a.out_score[(long long)b * a.max_det + k] = score[j];
  }
  if (t == 0) a.out_count[b] = nk;
}

}  // namespace b2p

using namespace b2p;

extern "C" {

int b2p_yolo_decode(const float* const* cls, const float* const* box, const int* Hs, const int* Ws, int nc, int B,
                    float conf, const float* pad_l, const float* pad_t, const float* scale, int cap, float* cand_box,
                    float* cand_score, int* cand_cls, int* cand_count, float* dense_ltrb, float* dense_score,
                    cudaStream_t st) {
  DecodeArgs a{};
  for (int s = 0; s < 3; ++s) { a.cls[s] = cls[s]; a.box[s] = box[s]; a.H[s] = Hs[s]; a.W[s] = Ws[s]; }
  a.nc = nc; a.B = B; a.cap = cap; a.conf = conf; a.pad_l = pad_l; a.pad_t = pad_t; a.scale = scale;
  a.cand_box = cand_box; a.cand_score = cand_score; a.cand_cls = cand_cls; a.cand_count = cand_count;
  a.dense_ltrb = dense_ltrb; a.dense_score = dense_score;
  yolo_decode_kernel<<<B, 1024, 0, st>>>(a);
  B2P_CHECK_LAUNCH();
  return 0;
}

int b2p_batched_nms(const float* box, const float* score, const int* cls, const int* count, int B, int cap,
                    double iou_thr, int max_det, const float* img_w, const float* img_h, int* keep_idx,
                    float* out_box, float* out_score, int* out_count, cudaStream_t st) {
  if (max_det > kNmsMaxKeep) return set_error("batched_nms: max_det > 1024 unsupported");
  if (cap > 16384) return set_error("batched_nms: candidate capacity > 16384 unsupported");
  NmsArgs a{};
  a.box = box; a.score = score; a.cls = cls; a.count = count; a.cap = cap; a.max_det = max_det;
  int P = 2;   // >= 2 keeps the float4 arrays carved after keys[] 16-byte aligned
  while (P < cap) P <<= 1;
  a.sort_cap = P;
  // torchvision compares (double)ovr > iou_threshold; for a float ovr that equals ovr > T with T the
  // largest float <= iou_threshold.
  float t = float(iou_thr);
  if (double(t) > iou_thr) t = nextafterf(t, -INFINITY);
  a.thr = t;
  a.img_w = img_w; a.img_h = img_h; a.keep_idx = keep_idx; a.out_box = out_box; a.out_score = out_score; a.out_count = out_count;
  const size_t dyn = size_t(P) * sizeof(unsigned long long) + kNmsMaxKeep * 28 + kNmsThreads * 24 + kNmsThreads * (kNmsThreads / 32) * 4;
  if (int e = bind_device()) return e;
  static std::atomic<bool> attr_set{false};   // idempotent attribute: a race between two first callers is harmless
  if (!attr_set.load()) {
    if (cudaFuncSetAttribute(batched_nms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8 + kNmsMaxKeep * 28 + kNmsThreads * 24 + kNmsThreads * (kNmsThreads / 32) * 4) != cudaSuccess)
      return set_error("batched_nms: cannot raise dynamic shared memory limit");
    attr_set = true;
  }
  batched_nms_kernel<<<B, kNmsThreads, dyn, st>>>(a);
  B2P_CHECK_LAUNCH();
  return 0;
}

}  // extern "C"

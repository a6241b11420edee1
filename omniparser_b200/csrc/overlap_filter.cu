// Overlap filter of get_som_labeled_img on the device: ref:util/utils.py:241-319 (remove_overlap_new),
// :411-415 (int_box_area), :432 (xyxy / whwh), :444-451 (element construction + "content is None last" sort).
//
// The reference does this arithmetic in Python floats (float64) on float32-rounded ratios, with strict comparisons and in
// list order.  The kernel keeps exactly that: ratios by IEEE fp32 division, everything after in fp64 through explicit
// round-to-nearest intrinsics (no FMA contraction), sums in the reference's association order.  Strings never come here:
// OCR boxes arrive as ratio boxes (already int_box_area-filtered on the host, where the texts live) and leave as
// "removed" flags plus, per icon, the bit mask of the OCR boxes whose text labels it.
//
// One CTA per screenshot (<= max_det icons x <= 32*mask_words OCR boxes); the last CTA to finish compacts the boxes that
// still need a caption (state 1) of all screenshots, screenshot-major, into the crop list b2p_crop_resize consumes.
#include "b2p_internal.h"

namespace b2p {

static constexpr int kOvlThreads = 256;
static constexpr int kOvlMaxIcons = 512;
static constexpr int kOvlMaxOcr = 512;

struct OvlArgs {
  const float* box_px; const int* count; int B, max_det;
  const float* img_w; const float* img_h;
  const float* ocr_ratio; const int* ocr_count; int max_ocr, mask_words;
  double thr;
  int* icon_state; unsigned* label_mask; int* ocr_removed;
  float* icon_ratio;                 // [B][max_det][4]: the fp32 ratio boxes (ref:util/utils.py:432), for the host element lists
  float* crop_box; int* crop_img; int* crop_counts;   // crop_counts [B + 1]: per screenshot, then the total
  int* arrive;                       // self-resetting arrival counter (last CTA compacts)
};

__device__ __forceinline__ double d_max0(double x) { return x > 0.0 ? x : 0.0; }   // Python max(0, x)

// intersection area of two xyxy boxes: max(0, min(x2) - max(x1)) * max(0, min(y2) - max(y1))
__device__ __forceinline__ double d_inter(const double* a, const double* b) {
  const double iw = d_max0(__dsub_rn(fmin(a[2], b[2]), fmax(a[0], b[0])));
  const double ih = d_max0(__dsub_rn(fmin(a[3], b[3]), fmax(a[1], b[1])));
  return __dmul_rn(iw, ih);
}

__global__ void __launch_bounds__(kOvlThreads) overlap_filter_kernel(OvlArgs a) {
  pdl_wait();
  __shared__ double ib[kOvlMaxIcons][4];      // icons that passed int_box_area, in order
  __shared__ double iarea[kOvlMaxIcons];
  __shared__ short isrc[kOvlMaxIcons];        // index into the detector's list
  __shared__ unsigned char ivalid[kOvlMaxIcons];
  __shared__ double ob[kOvlMaxOcr][4];
  __shared__ double oarea[kOvlMaxOcr];
  __shared__ unsigned orem[kOvlMaxOcr / 32];
  __shared__ int warp_cnt[kOvlThreads / 32 + 1];
  __shared__ int s_n, s_last;
  const int b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nraw = min(a.count[b], a.max_det);
  const int m = min(a.ocr_count[b], a.max_ocr);
  const float wf = a.img_w[b], hf = a.img_h[b];
  const double wd = double(wf), hd = double(hf);
  int* state = a.icon_state + (long long)b * a.max_det;
  unsigned* lmask = a.label_mask + (long long)b * a.max_det * a.mask_words;
  for (int i = tid; i < a.max_det; i += kOvlThreads) {
    state[i] = 0;
    for (int w = 0; w < a.mask_words; ++w) lmask[(long long)i * a.mask_words + w] = 0u;
  }
  if (tid < kOvlMaxOcr / 32) orem[tid] = 0u;
  if (tid == 0) s_n = 0;
  __syncthreads();
  // ---- 1. ratios (fp32 division, ref:util/utils.py:432), int_box_area > 0 (:411-415, :445), order-preserving compaction
  for (int base = 0; base < nraw; base += kOvlThreads) {
    const int i = base + tid;
    double r[4] = {0, 0, 0, 0};
    bool keep = false;
    if (i < nraw) {
      const float* p = a.box_px + ((long long)b * a.max_det + i) * 4;
      const float rf[4] = {__fdiv_rn(p[0], wf), __fdiv_rn(p[1], hf), __fdiv_rn(p[2], wf), __fdiv_rn(p[3], hf)};
      float* ro = a.icon_ratio + ((long long)b * a.max_det + i) * 4;
      for (int q = 0; q < 4; ++q) { r[q] = double(rf[q]); ro[q] = rf[q]; }
      const long long ia = ((long long)__dmul_rn(r[2], wd) - (long long)__dmul_rn(r[0], wd)) *
                           ((long long)__dmul_rn(r[3], hd) - (long long)__dmul_rn(r[1], hd));
      keep = ia > 0;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = s_n;
    for (int w = 0; w < warp; ++w) off += warp_cnt[w];
    if (keep) {
      const int d = off + __popc(bal & ((1u << lane) - 1u));
      for (int q = 0; q < 4; ++q) ib[d][q] = r[q];
      iarea[d] = __dmul_rn(__dsub_rn(r[2], r[0]), __dsub_rn(r[3], r[1]));
      isrc[d] = short(i);
    }
    __syncthreads();
    if (tid == 0) { int t = s_n; for (int w = 0; w < kOvlThreads / 32; ++w) t += warp_cnt[w]; s_n = t; }
    __syncthreads();
  }
  const int n = s_n;
  for (int k = tid; k < m; k += kOvlThreads) {
    const float* p = a.ocr_ratio + ((long long)b * a.max_ocr + k) * 4;
    for (int q = 0; q < 4; ++q) ob[k][q] = double(p[q]);
    oarea[k] = __dmul_rn(__dsub_rn(double(p[2]), double(p[0])), __dsub_rn(double(p[3]), double(p[1])));
  }
  __syncthreads();
  // ---- 2. icon vs icon (:277-285): invalid iff some other box overlaps (IoU* > thr) and is SMALLER
  for (int i = tid; i < n; i += kOvlThreads) {
    const double a1 = iarea[i];
    bool valid = true;
    for (int j = 0; j < n && valid; ++j) {
      if (j == i) continue;
      const double a2 = iarea[j];
      if (!(a1 > a2)) continue;                      // the area test is the cheap half of the conjunction
      const double inter = d_inter(ib[i], ib[j]);
      const double uni = __dadd_rn(__dsub_rn(__dadd_rn(a1, a2), inter), 1e-6);
      double r1 = 0.0, r2 = 0.0;
      if (a1 > 0.0 && a2 > 0.0) { r1 = __ddiv_rn(inter, a1); r2 = __ddiv_rn(inter, a2); }
      const double iou = fmax(fmax(__ddiv_rn(inter, uni), r1), r2);
      if (iou > a.thr) valid = false;
    }
    ivalid[i] = valid ? 1 : 0;
  }
  __syncthreads();
  // ---- 3. icon vs OCR (:286-316), per valid icon in OCR order: OCR box inside the icon -> label + removed; the first OCR
  // box that contains the icon (without being inside it) drops the icon and stops the walk
  for (int i = tid; i < n; i += kOvlThreads) {
    if (!ivalid[i]) continue;
    const double a1 = iarea[i];
    unsigned* lm = lmask + (long long)isrc[i] * a.mask_words;
    bool dropped = false, any = false;
    unsigned cur = 0u;
    for (int k = 0; k < m; ++k) {
      const double inter = d_inter(ob[k], ib[i]);
      if (__ddiv_rn(inter, oarea[k]) > 0.80) {           // is_inside(ocr, icon)
        cur |= 1u << (k & 31);
        any = true;
        atomicOr(&orem[k >> 5], 1u << (k & 31));
      } else if (__ddiv_rn(inter, a1) > 0.80) {          // is_inside(icon, ocr)
        dropped = true;
      }
      if ((k & 31) == 31 || k == m - 1 || dropped) {
        if (cur) lm[k >> 5] = cur;
        cur = 0u;
      }
      if (dropped) break;
    }
    state[isrc[i]] = dropped ? 0 : (any ? 2 : 1);
  }
  __syncthreads();
  for (int k = tid; k < a.max_ocr; k += kOvlThreads)
    a.ocr_removed[(long long)b * a.max_ocr + k] = (k < m) ? int((orem[k >> 5] >> (k & 31)) & 1u) : 0;
  // ---- 4. boxes that still need a caption, in order: per-screenshot count now, batch-wide compaction by the last CTA
  if (tid == 0) {
    int c = 0;
    for (int i = 0; i < nraw; ++i) c += (state[i] == 1);
    a.crop_counts[b] = c;
    __threadfence();
    s_last = (atomicAdd(a.arrive, 1) == a.B - 1);
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  if (tid == 0) {
    int t = 0;
    for (int q = 0; q < a.B; ++q) t += __ldcg(a.crop_counts + q);
    a.crop_counts[a.B] = t;
    *a.arrive = 0;
  }
  __syncthreads();
  int off = 0;
  for (int q = 0; q < a.B; ++q) {
    const int* st = a.icon_state + (long long)q * a.max_det;
    const int nq = min(__ldcg(a.count + q), a.max_det);
    // order-preserving compaction of this screenshot's state-1 boxes (block-wide, chunks of kOvlThreads)
    for (int base = 0; base < nq; base += kOvlThreads) {
      const int i = base + tid;
      const bool keep = (i < nq) && (__ldcg(st + i) == 1);
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      __syncthreads();
      if (lane == 0) warp_cnt[warp] = __popc(bal);
      __syncthreads();
      int o = off;
      for (int w = 0; w < warp; ++w) o += warp_cnt[w];
      if (keep) {
        const int d = o + __popc(bal & ((1u << lane) - 1u));
        const float* r = a.icon_ratio + ((long long)q * a.max_det + i) * 4;
        for (int c = 0; c < 4; ++c) a.crop_box[(long long)d * 4 + c] = __ldcg(r + c);
        a.crop_img[d] = q;
      }
      int tot = 0;
      for (int w = 0; w < kOvlThreads / 32; ++w) tot += warp_cnt[w];
      off += tot;
    }
  }
}

}  // namespace b2p

using namespace b2p;

extern "C" int b2p_overlap_filter(const float* box_px, const int* count, int B, int max_det, const float* img_w,
                                  const float* img_h, const float* ocr_ratio, const int* ocr_count, int max_ocr,
                                  double iou_thr, int* icon_state, unsigned* label_mask, int* ocr_removed, float* icon_ratio,
                                  float* crop_box, int* crop_img, int* crop_counts, int* arrive, cudaStream_t st) {
  if (int e = bind_device()) return e;
  if (B <= 0) return 0;
  if (max_det <= 0 || max_det > kOvlMaxIcons) return set_error("overlap_filter: max_det must be in 1..512");
  if (max_ocr <= 0 || max_ocr > kOvlMaxOcr || (max_ocr % 32) != 0) return set_error("overlap_filter: max_ocr must be a multiple of 32 in 32..512");
  OvlArgs a{box_px, count, B, max_det, img_w, img_h, ocr_ratio, ocr_count, max_ocr, max_ocr / 32, iou_thr,
            icon_state, label_mask, ocr_removed, icon_ratio, crop_box, crop_img, crop_counts, arrive};
  overlap_filter_kernel<<<B, kOvlThreads, 0, st>>>(a);
  B2P_CHECK_LAUNCH();
  return 0;
}

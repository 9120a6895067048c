// Stage 3b: brute-force Hamming 2-NN + Lowe ratio, batched RANSAC over 4-point hypotheses (one warp each) with
// local optimisation, and the box warp.
//
// Restates, as stabilo calls them (/root/reference/geotrax/extract.py:181-187):
//   cv2.BFMatcher(NORM_HAMMING).knnMatch(q, t, k=2) + `m.distance < ratio * n.distance`   (bit-exact, ties -> lower index)
//   cv2.findHomography(cur, ref, USAC_MAGSAC, 2.0, maxIters=5000)  (own estimator; geometric parity)
//   Stabilizer.transform_cur_boxes(): 4 corners -> H -> axis-aligned envelope -> xywh   (pinned by the golden files)
#include <cmath>

#include "engine.cuh"

namespace {

// =====================================================================================================================
// Hamming 2-NN.  Block = 8 warps x 4 queries; train descriptors staged through shared memory in word-major tiles.
// =====================================================================================================================
constexpr int kTile = 256;
constexpr int kQPW = 4;

__global__ void __launch_bounds__(256) match_kernel(const uint8_t* __restrict__ qdesc, size_t q_stride, const int* __restrict__ nq, int nq_step,
                                                    const uint8_t* __restrict__ tdesc, size_t t_stride, const int* __restrict__ nt, int nt_step,
                                                    int* __restrict__ out_idx, int* __restrict__ out_dist, size_t out_stride) {
  __shared__ unsigned s_t[8][kTile];
  const int b = blockIdx.y;
  const int NQ = min(nq[b * nq_step], GT_MAX_KP), NT = min(nt[b * nt_step], GT_MAX_KP);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = (blockIdx.x * 8 + warp) * kQPW;
  if (blockIdx.x * 8 * kQPW >= NQ) return;
  const uint8_t* Q = qdesc + (size_t)b * q_stride;
  const uint8_t* T = tdesc + (size_t)b * t_stride;
  unsigned qw[kQPW][8];
#pragma unroll
  for (int i = 0; i < kQPW; ++i) {
    const int q = min(q0 + i, NQ - 1);
    const uint4 a = __ldg(reinterpret_cast<const uint4*>(Q + (size_t)q * 32));
    const uint4 c = __ldg(reinterpret_cast<const uint4*>(Q + (size_t)q * 32) + 1);
    qw[i][0] = a.x; qw[i][1] = a.y; qw[i][2] = a.z; qw[i][3] = a.w; qw[i][4] = c.x; qw[i][5] = c.y; qw[i][6] = c.z; qw[i][7] = c.w;
  }
  unsigned k1[kQPW], k2[kQPW];  // (distance << 16 | index), smaller is better
#pragma unroll
  for (int i = 0; i < kQPW; ++i) k1[i] = k2[i] = 0xFFFFFFFFu;
  for (int t0 = 0; t0 < NT; t0 += kTile) {
    __syncthreads();
    for (int i = threadIdx.x; i < kTile * 2; i += 256) {
      const int j = i >> 1, half = i & 1;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (t0 + j < NT) v = __ldg(reinterpret_cast<const uint4*>(T + (size_t)(t0 + j) * 32) + half);
      s_t[half * 4 + 0][j] = v.x; s_t[half * 4 + 1][j] = v.y; s_t[half * 4 + 2][j] = v.z; s_t[half * 4 + 3][j] = v.w;
    }
    __syncthreads();
    const int lim = min(kTile, NT - t0);
    for (int j = lane; j < lim; j += 32) {
      unsigned tw[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) tw[k] = s_t[k][j];
#pragma unroll
      for (int i = 0; i < kQPW; ++i) {
        int d = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) d += __popc(qw[i][k] ^ tw[k]);
        const unsigned key = ((unsigned)d << 16) | (unsigned)(t0 + j);
        if (key < k1[i]) { k2[i] = k1[i]; k1[i] = key; }
        else if (key < k2[i]) k2[i] = key;
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kQPW; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned o1 = __shfl_xor_sync(0xffffffffu, k1[i], o), o2 = __shfl_xor_sync(0xffffffffu, k2[i], o);
      const unsigned lo = min(k1[i], o1), hi = max(k1[i], o1);
      k2[i] = min(hi, min(k2[i], o2));
      k1[i] = lo;
    }
    const int q = q0 + i;
    if (lane == 0 && q < NQ) {
      int* oi = out_idx + (size_t)b * out_stride + (size_t)q * 2;
      int* od = out_dist + (size_t)b * out_stride + (size_t)q * 2;
      oi[0] = k1[i] == 0xFFFFFFFFu ? -1 : (int)(k1[i] & 0xFFFF); od[0] = k1[i] == 0xFFFFFFFFu ? -1 : (int)(k1[i] >> 16);
      oi[1] = k2[i] == 0xFFFFFFFFu ? -1 : (int)(k2[i] & 0xFFFF); od[1] = k2[i] == 0xFFFFFFFFu ? -1 : (int)(k2[i] >> 16);
    }
  }
}

// ---- ratio test + ordered compaction into point pairs (cur x,y | ref x,y), one block per frame --------------------------------
__global__ void __launch_bounds__(1024) build_pairs_kernel(const int* __restrict__ midx, const int* __restrict__ mdist, const float* __restrict__ kp_all,
                                                           const int* __restrict__ kp_count, int ref_slot, int query_is_current, double ratio,
                                                           float pt_scale, float* __restrict__ pairs, int* __restrict__ pair_count) {
  __shared__ int s_warp[32];
  __shared__ int s_base;
  const int b = blockIdx.x;
  const int nq = min(query_is_current ? kp_count[b] : kp_count[ref_slot], GT_MAX_KP);
  const int* mi = midx + (size_t)b * GT_MAX_KP * 2;
  const int* md = mdist + (size_t)b * GT_MAX_KP * 2;
  const float* kq = kp_all + (size_t)(query_is_current ? b : ref_slot) * GT_MAX_KP * 6;
  const float* kt = kp_all + (size_t)(query_is_current ? ref_slot : b) * GT_MAX_KP * 6;
  float* out = pairs + (size_t)b * GT_MAX_KP * 4;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q0 = 0; q0 < nq; q0 += blockDim.x) {
    const int q = q0 + threadIdx.x;
    bool good = false;
    if (q < nq && mi[q * 2 + 1] >= 0) good = (double)md[q * 2] < ratio * (double)md[q * 2 + 1];
    const unsigned bal = __ballot_sync(0xffffffffu, good);
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    int wbase = s_base;
    for (int w = 0; w < warp; ++w) wbase += s_warp[w];
    if (good) {
      const int pos = wbase + __popc(bal & ((1u << lane) - 1));
      const float* a = kq + (size_t)q * 6;
      const float* c = kt + (size_t)mi[q * 2] * 6;
      const float* cur = query_is_current ? a : c;
      const float* ref = query_is_current ? c : a;
      out[pos * 4 + 0] = cur[0] * pt_scale; out[pos * 4 + 1] = cur[1] * pt_scale;
      out[pos * 4 + 2] = ref[0] * pt_scale; out[pos * 4 + 3] = ref[1] * pt_scale;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += s_warp[w];
      s_base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) pair_count[b] = s_base;
}

// =====================================================================================================================
// Robust homography
// =====================================================================================================================
struct Norm { float mcx, mcy, sc, mrx, mry, sr; };

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int N>
__device__ void block_sum(double* v, double* s_buf /* [N * 8] */, double* s_out /* [N] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < N; ++i) {
    const double r = warp_sum_d(v[i]);
    if (lane == 0) s_buf[i * 8 + warp] = r;
  }
  __syncthreads();
  if (threadIdx.x < N) {
    double a = 0;
    for (int w = 0; w < 8; ++w) a += s_buf[threadIdx.x * 8 + w];
    s_out[threadIdx.x] = a;
  }
  __syncthreads();
}

// Hartley normalisation of both point sets; writes normalised pairs
__global__ void __launch_bounds__(256) ransac_prepare_kernel(const float* __restrict__ pairs, const int* __restrict__ counts, int pair_stride,
                                                             float4* __restrict__ npairs, Norm* __restrict__ norms) {
  __shared__ double s_buf[6 * 8], s_out[6];
  const int b = blockIdx.x;
  const int m = min(counts[b], pair_stride);
  const float4* p = reinterpret_cast<const float4*>(pairs + (size_t)b * pair_stride * 4);
  double v[6] = {0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < m; i += 256) {
    const float4 q = p[i];
    v[0] += q.x; v[1] += q.y; v[2] += q.z; v[3] += q.w;
  }
  block_sum<4>(v, s_buf, s_out);
  const double inv = m > 0 ? 1.0 / m : 0.0;
  const double mcx = s_out[0] * inv, mcy = s_out[1] * inv, mrx = s_out[2] * inv, mry = s_out[3] * inv;
  double d[2] = {0, 0};
  for (int i = threadIdx.x; i < m; i += 256) {
    const float4 q = p[i];
    d[0] += sqrt((q.x - mcx) * (q.x - mcx) + (q.y - mcy) * (q.y - mcy));
    d[1] += sqrt((q.z - mrx) * (q.z - mrx) + (q.w - mry) * (q.w - mry));
  }
  block_sum<2>(d, s_buf, s_out);
  const double sc = s_out[0] > 0 ? 1.4142135623730951 * m / s_out[0] : 1.0;
  const double sr = s_out[1] > 0 ? 1.4142135623730951 * m / s_out[1] : 1.0;
  float4* o = npairs + (size_t)b * pair_stride;
  for (int i = threadIdx.x; i < m; i += 256) {
    const float4 q = p[i];
    o[i] = make_float4((float)((q.x - mcx) * sc), (float)((q.y - mcy) * sc), (float)((q.z - mrx) * sr), (float)((q.w - mry) * sr));
  }
  if (threadIdx.x == 0) {
    Norm n;
    n.mcx = (float)mcx; n.mcy = (float)mcy; n.sc = (float)sc; n.mrx = (float)mrx; n.mry = (float)mry; n.sr = (float)sr;
    norms[b] = n;
  }
}

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

// deterministic 4-sample for hypothesis h; returns false if 4 distinct indices could not be drawn.  The sample is a function of
// (seed, h, m) only -- NOT of the batch slot b the frame happens to sit in -- so a frame yields the same homography whichever batch
// composition / rank / shard processes it (the sharded flight equals the single-GPU flight bit for bit, tests/test_gpu_round2.py).
__device__ __forceinline__ bool draw_sample(unsigned seed, int /*b*/, int h, int m, int* s) {
  unsigned st = hash32(seed ^ hash32(0x9E3779B9u + (unsigned)h));
  int got = 0;
  for (int tries = 0; tries < 16 && got < 4; ++tries) {
    st = hash32(st + 0x6D2B79F5u);
    const int c = (int)(st % (unsigned)m);
    bool dup = false;
    for (int j = 0; j < got; ++j) dup |= (s[j] == c);
    if (!dup) s[got++] = c;
  }
  return got == 4;
}

// columns (x,y,1) of three points scaled so that their sum maps to the fourth: A = [p1 p2 p3] diag(adj([p1 p2 p3]) p4)
__device__ __forceinline__ bool proj_basis(const float* px, const float* py, float* M) {
  const float x1 = px[0], y1 = py[0], x2 = px[1], y2 = py[1], x3 = px[2], y3 = py[2], x4 = px[3], y4 = py[3];
  // adj(M) * p4 with M = [[x1,x2,x3],[y1,y2,y3],[1,1,1]]
  const float l1 = (y2 - y3) * x4 + (x3 - x2) * y4 + (x2 * y3 - x3 * y2);
  const float l2 = (y3 - y1) * x4 + (x1 - x3) * y4 + (x3 * y1 - x1 * y3);
  const float l3 = (y1 - y2) * x4 + (x2 - x1) * y4 + (x1 * y2 - x2 * y1);
  const float det = x1 * (y2 - y3) - x2 * (y1 - y3) + x3 * (y1 - y2);
  const float eps = 1e-6f;
  if (fabsf(l1) < eps || fabsf(l2) < eps || fabsf(l3) < eps || fabsf(det) < eps) return false;
  M[0] = x1 * l1; M[1] = x2 * l2; M[2] = x3 * l3;
  M[3] = y1 * l1; M[4] = y2 * l2; M[5] = y3 * l3;
  M[6] = l1;      M[7] = l2;      M[8] = l3;
  return true;
}

__device__ __forceinline__ bool homography_4pt(const float4* __restrict__ np, const int* s, float* H) {
  float cx[4], cy[4], rx[4], ry[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 q = np[s[k]];
    cx[k] = q.x; cy[k] = q.y; rx[k] = q.z; ry[k] = q.w;
  }
  float A[9], Bm[9];
  if (!proj_basis(cx, cy, A) || !proj_basis(rx, ry, Bm)) return false;
  // H = B * adj(A)
  float J[9];
  J[0] = A[4] * A[8] - A[5] * A[7]; J[1] = A[2] * A[7] - A[1] * A[8]; J[2] = A[1] * A[5] - A[2] * A[4];
  J[3] = A[5] * A[6] - A[3] * A[8]; J[4] = A[0] * A[8] - A[2] * A[6]; J[5] = A[2] * A[3] - A[0] * A[5];
  J[6] = A[3] * A[7] - A[4] * A[6]; J[7] = A[1] * A[6] - A[0] * A[7]; J[8] = A[0] * A[4] - A[1] * A[3];
  float nrm = 0.f;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = Bm[r * 3 + 0] * J[0 * 3 + c] + Bm[r * 3 + 1] * J[1 * 3 + c] + Bm[r * 3 + 2] * J[2 * 3 + c];
      H[r * 3 + c] = v;
      nrm += v * v;
    }
  if (!(nrm > 1e-30f) || !isfinite(nrm)) return false;
  float inv = rsqrtf(nrm);
  // orientation: w must keep one sign on the four sample points; make it positive
  float wmin = 1e30f, wmax = -1e30f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float w = H[6] * cx[k] + H[7] * cy[k] + H[8];
    wmin = fminf(wmin, w);
    wmax = fmaxf(wmax, w);
  }
  if (wmin * wmax <= 0.f) return false;
  if (wmax < 0.f) inv = -inv;
#pragma unroll
  for (int i = 0; i < 9; ++i) H[i] *= inv;
  return true;
}

__device__ __forceinline__ float reproj_err2(const float* H, const float4 q) {
  const float w = H[6] * q.x + H[7] * q.y + H[8];
  if (!(w > 1e-8f)) return 1e30f;
  float iw;   // approximate reciprocal (1 MUFU, ~1 ulp): the score is a sum of soft inlier weights, its last bits do not matter
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iw) : "f"(w));
  const float du = (H[0] * q.x + H[1] * q.y + H[2]) * iw - q.z;
  const float dv = (H[3] * q.x + H[4] * q.y + H[5]) * iw - q.w;
  return du * du + dv * dv;
}

// MSAC score of a hypothesis = sum over matches of max(0, 1 - e^2 / t^2).
//
// Preemptive scoring in two launches (the exhaustive round-1 kernel -- one warp per hypothesis, 5,000 hypotheses x ~1,600 matches x 16
// frames -- was instruction bound at 155-205 us):
//   pass 1  ransac_score_kernel: ONE THREAD per hypothesis scores it on the strided subset {0, sub, 2 sub, ...} of <= kSubsetPairs matches
//           (matches are ordered by key-point index = pyramid level, then position, so a stride covers every level and the whole image;
//           all threads of a block read the same match at the same time = shared-memory broadcast) and files the score in a
//           per-frame 256-bin histogram (bin = score / subset size: the score of a hypothesis cannot exceed the number of matches);
//   pass 2  ransac_rescore_kernel: walks the histogram from the top to the bin where kKeepHyp hypotheses are covered, scores the
//           hypotheses at or above that bin on ALL matches (one warp each) and rewrites the score array: full score for the survivors,
//           -1 for everything else -- ransac_finalize_kernel's arg-max (ties -> lowest hypothesis index) runs unchanged.
// 6x less scoring work; the winner is the survivor with the best FULL score, and the local optimisation starts from it as before.
// Everything is a deterministic function of the frame's matches (histogram counts, not arrival order, decide who survives).
constexpr int kScoreThreads = 256, kSubsetPairs = 256, kKeepHyp = 64;
__global__ void __launch_bounds__(kScoreThreads) ransac_score_kernel(const float4* __restrict__ npairs, const int* __restrict__ counts, int pair_stride,
                                                                     const Norm* __restrict__ norms, float thr, int max_iter, unsigned seed,
                                                                     float* __restrict__ scores, int* __restrict__ hist) {
  __shared__ float4 s_np[kSubsetPairs * 2];
  const int b = blockIdx.y;
  const int h = blockIdx.x * kScoreThreads + threadIdx.x;
  const int m = min(counts[b], pair_stride);
  const int sub = max(1, m / kSubsetPairs);          // pass 1 looks at every sub-th match: kSubsetPairs <= ms < 2 kSubsetPairs of them
  const int ms = (m + sub - 1) / sub;
  const float4* np = npairs + (size_t)b * pair_stride;
  for (int i = threadIdx.x; i < ms; i += kScoreThreads) s_np[i] = __ldg(&np[(size_t)i * sub]);
  __syncthreads();
  if (h >= max_iter) return;
  int s[4];
  float H[9];
  const bool ok = m >= 4 && draw_sample(seed, b, h, m, s) && homography_4pt(np, s, H);
  float a = -1.f;
  if (ok) {
    const float t = thr * norms[b].sr;
    const float it2 = 1.0f / (t * t);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    int i = 0;
    for (; i + 3 < ms; i += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] += fmaxf(0.f, 1.0f - reproj_err2(H, s_np[i + u]) * it2);
    }
    for (; i < ms; ++i) acc[0] += fmaxf(0.f, 1.0f - reproj_err2(H, s_np[i]) * it2);
    a = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    atomicAdd(&hist[b * 256 + min(255, (int)(a * 255.0f / (float)ms))], 1);
  }
  scores[(size_t)b * max_iter + h] = a;
}

// One block per frame, 32 warps.  Survivors = every hypothesis in a bin above the cut bin + the lowest-index hypotheses of the cut bin
// up to kKeepHyp in total (an ordered compaction: which hypotheses survive depends on their index, never on thread timing).
__global__ void __launch_bounds__(1024) ransac_rescore_kernel(const float4* __restrict__ npairs, const int* __restrict__ counts, int pair_stride,
                                                              const Norm* __restrict__ norms, float thr, int max_iter, unsigned seed,
                                                              float* __restrict__ scores, const int* __restrict__ hist) {
  __shared__ int s_cut, s_need, s_nhi, s_ncut;
  __shared__ int s_whi[32], s_wcut[32];
  __shared__ int s_list[kKeepHyp];
  __shared__ float s_full[kKeepHyp];
  const int b = blockIdx.x;
  const int m = min(counts[b], pair_stride);
  const int sub = max(1, m / kSubsetPairs);
  if (m < 4 || sub == 1) return;                             // pass 1 already scored every match
  const int ms = (m + sub - 1) / sub;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {      // cut bin = the highest bin q with count(bins >= q) >= kKeepHyp; need = how many of bin q's members complete the set
    int mine = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) mine += hist[b * 256 + 255 - 8 * lane - j];      // lane l owns bins 255 - 8 l .. 248 - 8 l (descending)
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    const unsigned hit = __ballot_sync(0xffffffffu, incl >= kKeepHyp);
    int cut = -1, need = 0;                                  // fewer than kKeepHyp valid hypotheses: all of them survive (cut -1)
    if (hit) {
      const int src = __ffs(hit) - 1;
      if (lane == src) {
        int acc = incl - mine;
        for (int j = 0; j < 8; ++j) {
          const int q = 255 - 8 * lane - j, c = hist[b * 256 + q];
          if (acc + c >= kKeepHyp) { cut = q; need = kKeepHyp - acc; break; }
          acc += c;
        }
      }
      cut = __shfl_sync(0xffffffffu, cut, src);
      need = __shfl_sync(0xffffffffu, need, src);
    }
    if (lane == 0) { s_cut = cut; s_need = need; s_nhi = 0; s_ncut = 0; }
  }
  __syncthreads();
  float* sc = scores + (size_t)b * max_iter;
  const int cut = s_cut, need = s_need;
  for (int h0 = 0; h0 < max_iter; h0 += 1024) {
    const int h = h0 + threadIdx.x;
    const float part = h < max_iter ? sc[h] : -1.f;
    const int bin = part >= 0.f ? min(255, (int)(part * 255.0f / (float)ms)) : -2;
    const bool hi = bin > cut, eq = bin == cut && bin >= 0;
    const unsigned bhi = __ballot_sync(0xffffffffu, hi), beq = __ballot_sync(0xffffffffu, eq);
    if (lane == 0) { s_whi[warp] = __popc(bhi); s_wcut[warp] = __popc(beq); }
    __syncthreads();
    int phi = s_nhi, pcut = s_ncut;
    for (int w = 0; w < warp; ++w) { phi += s_whi[w]; pcut += s_wcut[w]; }
    // survivors are listed as [bins above the cut ... | cut-bin members ...]: slot = rank among "hi", or (kKeepHyp - need) + rank among "eq"
    if (hi) { const int pos = phi + __popc(bhi & ((1u << lane) - 1)); if (pos < kKeepHyp) s_list[pos] = h; }
    if (eq) { const int r = pcut + __popc(beq & ((1u << lane) - 1)); if (r < need) s_list[kKeepHyp - need + r] = h; }
    __syncthreads();
    if (threadIdx.x == 0) { int a = 0, c = 0; for (int w = 0; w < 32; ++w) { a += s_whi[w]; c += s_wcut[w]; } s_nhi += a; s_ncut += c; }
    __syncthreads();
  }
  // cut >= 0: exactly kKeepHyp survivors (nhi = kKeepHyp - need above the cut, `need` from the cut bin); cut == -1: the nhi valid ones
  const int nsel = cut >= 0 ? kKeepHyp : min(s_nhi, kKeepHyp);
  const float4* np = npairs + (size_t)b * pair_stride;
  const float t = thr * norms[b].sr;
  const float it2 = 1.0f / (t * t);
  for (int k = warp; k < nsel; k += 32) {
    int s[4];
    float H[9];
    const bool ok = draw_sample(seed, b, s_list[k], m, s) && homography_4pt(np, s, H);
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (ok) {
      int i = lane;
      for (; i + 96 < m; i += 128) {
#pragma unroll
        for (int u = 0; u < 4; ++u) acc[u] += fmaxf(0.f, 1.0f - reproj_err2(H, __ldg(&np[i + 32 * u])) * it2);
      }
      for (; i < m; i += 32) acc[0] += fmaxf(0.f, 1.0f - reproj_err2(H, __ldg(&np[i])) * it2);
    }
    float a = (acc[0] + acc[1]) + (acc[2] + acc[3]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if (lane == 0) s_full[k] = ok ? a : -1.f;
  }
  __syncthreads();
  for (int h = threadIdx.x; h < max_iter; h += 1024) sc[h] = -1.f;
  __syncthreads();
  for (int k = threadIdx.x; k < nsel; k += 1024) sc[s_list[k]] = s_full[k];
}

// 8x8 SPD solve (Cholesky, in place).  A is the full symmetric matrix row-major; returns false if not positive definite.
__device__ bool chol_solve8(double* A, double* rhs) {
  for (int j = 0; j < 8; ++j) {
    double d = A[j * 8 + j];
    for (int k = 0; k < j; ++k) d -= A[j * 8 + k] * A[j * 8 + k];
    if (!(d > 1e-300)) return false;
    d = sqrt(d);
    const double id = 1.0 / d;       // the solve runs on ONE thread: 8 divisions instead of 52
    A[j * 8 + j] = id;               // the diagonal stores 1 / L_jj
    for (int i = j + 1; i < 8; ++i) {
      double v = A[i * 8 + j];
      for (int k = 0; k < j; ++k) v -= A[i * 8 + k] * A[j * 8 + k];
      A[i * 8 + j] = v * id;
    }
  }
  for (int i = 0; i < 8; ++i) {
    double v = rhs[i];
    for (int k = 0; k < i; ++k) v -= A[i * 8 + k] * rhs[k];
    rhs[i] = v * A[i * 8 + i];
  }
  for (int i = 7; i >= 0; --i) {
    double v = rhs[i];
    for (int k = i + 1; k < 8; ++k) v -= A[k * 8 + i] * rhs[k];
    rhs[i] = v * A[i * 8 + i];
  }
  return true;
}

// pick the best hypothesis, then local optimisation: weighted linear fit -> Gauss-Newton on the reprojection error,
// re-selecting inliers between rounds; finally de-normalise / conjugate to source-frame pixels.
__global__ void __launch_bounds__(256) ransac_finalize_kernel(const float4* __restrict__ npairs, const int* __restrict__ counts, int pair_stride,
                                                              const Norm* __restrict__ norms, const float* __restrict__ scores, float thr,
                                                              int max_iter, unsigned seed, float ratio, int full_res, double* __restrict__ out_H,
                                                              int* __restrict__ out_status, int* __restrict__ out_stats, int kp_ref_slot,
                                                              const int* __restrict__ kp_count) {
  __shared__ double s_buf[44 * 8], s_out[44];
  __shared__ double s_h[9];
  __shared__ float s_bs[8];
  __shared__ int s_bi[8];
  __shared__ int s_ok;
  const int b = blockIdx.x;
  const int m = min(counts[b], pair_stride);
  const float4* np = npairs + (size_t)b * pair_stride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int* stats = out_stats + b * 4;
  if (threadIdx.x == 0) {
    stats[0] = kp_count ? kp_count[kp_ref_slot] : 0;
    stats[1] = kp_count ? kp_count[b] : 0;
    stats[2] = m;
    stats[3] = 0;
  }
  auto fail = [&]() {
    if (threadIdx.x == 0) {
      out_status[b] = 1;
      for (int i = 0; i < 9; ++i) out_H[b * 9 + i] = (i % 4 == 0) ? 1.0 : 0.0;
    }
  };
  if (m < 4) { fail(); return; }
  // argmax (ties -> lowest hypothesis index)
  float bs = -2.f;
  int bi = 0x7fffffff;
  for (int h = threadIdx.x; h < max_iter; h += 256) {
    const float s = scores[(size_t)b * max_iter + h];
    if (s > bs || (s == bs && h < bi)) { bs = s; bi = h; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float os = __shfl_xor_sync(0xffffffffu, bs, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (os > bs || (os == bs && oi < bi)) { bs = os; bi = oi; }
  }
  if (lane == 0) { s_bs[warp] = bs; s_bi[warp] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w)
      if (s_bs[w] > s_bs[0] || (s_bs[w] == s_bs[0] && s_bi[w] < s_bi[0])) { s_bs[0] = s_bs[w]; s_bi[0] = s_bi[w]; }
    int s[4];
    float H[9];
    s_ok = (s_bs[0] > 0.f) && draw_sample(seed, b, s_bi[0], m, s) && homography_4pt(np, s, H) && fabsf(H[8]) > 1e-12f;
    if (s_ok)
      for (int i = 0; i < 9; ++i) s_h[i] = (double)H[i] / (double)H[8];
  }
  __syncthreads();
  if (!s_ok) { fail(); return; }
  const double t = (double)thr * (double)norms[b].sr;
  const double t2 = t * t;
  // Local optimisation.  Round 0: linear (algebraic, h33 = 1) fit on the inliers of the best sample; rounds 1..5:
  // Gauss-Newton steps on the forward reprojection error with uniform weights, re-selecting the support each round.
  // From round 2 on the support is every match within 2x the threshold: with integer-pixel ORB positions the inlier
  // residuals are quantisation-dominated (bimodal), and a wide uniform support is what keeps the fit unbiased -- this
  // is also where MAGSAC++'s sigma-marginalised weights put their mass (measured: centres agree with cv2 USAC_MAGSAC
  // to 0.06 px mean on the synthetic flights, vs 0.19 px with a 1x support and 0.6 px with Tukey weights).
  for (int round = 0; round < 6; ++round) {
    const double tsel = round >= 2 ? 4.0 * t2 : t2;
    double h[9];
    for (int i = 0; i < 9; ++i) h[i] = s_h[i];
    double acc[44];
    for (int i = 0; i < 44; ++i) acc[i] = 0;
    for (int i = threadIdx.x; i < m; i += 256) {
      const float4 q = np[i];
      const double x = q.x, y = q.y, u = q.z, v = q.w;
      const double w = h[6] * x + h[7] * y + h[8];
      if (!(w > 1e-9)) continue;
      const double iw = 1.0 / w;   // one double division per match instead of three
      const double pu = (h[0] * x + h[1] * y + h[2]) * iw, pv = (h[3] * x + h[4] * y + h[5]) * iw;
      const double e2 = (pu - u) * (pu - u) + (pv - v) * (pv - v);
      if (!(e2 < tsel)) continue;
      const double wt = 1.0;
      double r0[8], r1[8], b0, b1;
      if (round == 0) {
        r0[0] = x; r0[1] = y; r0[2] = 1; r0[3] = 0; r0[4] = 0; r0[5] = 0; r0[6] = -u * x; r0[7] = -u * y; b0 = u;
        r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = x; r1[4] = y; r1[5] = 1; r1[6] = -v * x; r1[7] = -v * y; b1 = v;
      } else {
        r0[0] = x * iw; r0[1] = y * iw; r0[2] = iw; r0[3] = 0; r0[4] = 0; r0[5] = 0; r0[6] = -pu * x * iw; r0[7] = -pu * y * iw; b0 = u - pu;
        r1[0] = 0; r1[1] = 0; r1[2] = 0; r1[3] = x * iw; r1[4] = y * iw; r1[5] = iw; r1[6] = -pv * x * iw; r1[7] = -pv * y * iw; b1 = v - pv;
      }
      int k = 0;
#pragma unroll
      for (int a = 0; a < 8; ++a) {
#pragma unroll
        for (int c = a; c < 8; ++c) acc[k++] += wt * (r0[a] * r0[c] + r1[a] * r1[c]);
      }
#pragma unroll
      for (int a = 0; a < 8; ++a) acc[36 + a] += wt * (r0[a] * b0 + r1[a] * b1);
    }
    block_sum<44>(acc, s_buf, s_out);
    if (threadIdx.x == 0) {
      double A[64], rhs[8];
      int k = 0;
      double tr = 0;
      for (int a = 0; a < 8; ++a)
        for (int c = a; c < 8; ++c) { A[a * 8 + c] = s_out[k]; A[c * 8 + a] = s_out[k]; ++k; }
      for (int a = 0; a < 8; ++a) { rhs[a] = s_out[36 + a]; tr += A[a * 8 + a]; }
      for (int a = 0; a < 8; ++a) A[a * 8 + a] += 1e-12 * tr;
      if (chol_solve8(A, rhs)) {
        if (round == 0) { for (int a = 0; a < 8; ++a) s_h[a] = rhs[a]; s_h[8] = 1.0; }
        else for (int a = 0; a < 8; ++a) s_h[a] += rhs[a];
      }
    }
    __syncthreads();
  }
  // inlier count at the nominal threshold
  {
    double h[9];
    for (int i = 0; i < 9; ++i) h[i] = s_h[i];
    double cnt[1] = {0};
    for (int i = threadIdx.x; i < m; i += 256) {
      const float4 q = np[i];
      const double w = h[6] * q.x + h[7] * q.y + h[8];
      if (!(w > 1e-9)) continue;
      const double iw = 1.0 / w;
      const double pu = (h[0] * q.x + h[1] * q.y + h[2]) * iw, pv = (h[3] * q.x + h[4] * q.y + h[5]) * iw;
      if ((pu - q.z) * (pu - q.z) + (pv - q.w) * (pv - q.w) < t2) cnt[0] += 1.0;
    }
    block_sum<1>(cnt, s_buf, s_out);
  }
  if (threadIdx.x == 0) {
    const Norm n = norms[b];
    const int inl = (int)(s_out[0] + 0.5);
    stats[3] = inl;
    // H = Tr^-1 * Hn * Tc,  Tc = [[sc,0,-sc*mcx],[0,sc,-sc*mcy],[0,0,1]],  Tr^-1 = [[1/sr,0,mrx],[0,1/sr,mry],[0,0,1]]
    const double* g = s_h;
    const double sc = n.sc, sr = n.sr;
    double M[9];  // Hn * Tc
    for (int r = 0; r < 3; ++r) {
      M[r * 3 + 0] = g[r * 3 + 0] * sc;
      M[r * 3 + 1] = g[r * 3 + 1] * sc;
      M[r * 3 + 2] = -g[r * 3 + 0] * sc * n.mcx - g[r * 3 + 1] * sc * n.mcy + g[r * 3 + 2];
    }
    double H[9];
    for (int c = 0; c < 3; ++c) {
      H[0 * 3 + c] = M[0 * 3 + c] / sr + n.mrx * M[2 * 3 + c];
      H[1 * 3 + c] = M[1 * 3 + c] / sr + n.mry * M[2 * 3 + c];
      H[2 * 3 + c] = M[2 * 3 + c];
    }
    if (!full_res && ratio != 1.0f) {  // conjugate working-resolution H to source-frame pixels: S^-1 H S, S = diag(r, r, 1)
      H[2] /= ratio; H[5] /= ratio; H[6] *= ratio; H[7] *= ratio;
    }
    const double det = H[0] * (H[4] * H[8] - H[5] * H[7]) - H[1] * (H[3] * H[8] - H[5] * H[6]) + H[2] * (H[3] * H[7] - H[4] * H[6]);
    bool ok = inl >= 4 && fabs(H[8]) > 1e-12 && isfinite(det);
    if (ok) {
      const double i33 = 1.0 / H[8];
      for (int i = 0; i < 9; ++i) out_H[b * 9 + i] = H[i] * i33;
      out_status[b] = 0;
    } else {
      for (int i = 0; i < 9; ++i) out_H[b * 9 + i] = (i % 4 == 0) ? 1.0 : 0.0;
      out_status[b] = 1;
    }
  }
}

// ---- box warp: 4 corners -> H -> axis-aligned envelope -> xywh ------------------------------------------------------------------
__global__ void warp_boxes_kernel(const double* __restrict__ Hs, const int* __restrict__ status, const float* __restrict__ in, float* __restrict__ out,
                                  const int* __restrict__ counts, int stride) {
  const int b = blockIdx.y;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= min(counts[b], stride)) return;
  const float* s = in + ((size_t)b * stride + i) * 4;
  float* o = out + ((size_t)b * stride + i) * 4;
  if (status && status[b] != 0) { o[0] = s[0]; o[1] = s[1]; o[2] = s[2]; o[3] = s[3]; return; }
  const double* H = Hs + (size_t)b * 9;
  const double x0 = (double)s[0] - (double)s[2] / 2, x1 = (double)s[0] + (double)s[2] / 2;
  const double y0 = (double)s[1] - (double)s[3] / 2, y1 = (double)s[1] + (double)s[3] / 2;
  const double cx[4] = {x0, x1, x1, x0}, cy[4] = {y0, y0, y1, y1};
  double umin = 1e300, umax = -1e300, vmin = 1e300, vmax = -1e300;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double w = H[6] * cx[k] + H[7] * cy[k] + H[8];
    const double u = (H[0] * cx[k] + H[1] * cy[k] + H[2]) / w, v = (H[3] * cx[k] + H[4] * cy[k] + H[5]) / w;
    umin = fmin(umin, u); umax = fmax(umax, u); vmin = fmin(vmin, v); vmax = fmax(vmax, v);
  }
  o[0] = (float)((umin + umax) / 2); o[1] = (float)((vmin + vmax) / 2); o[2] = (float)(umax - umin); o[3] = (float)(vmax - vmin);
}

}  // namespace

int stab_build(gt_engine* e) {
  const int B = e->cfg.max_batch;
  GT_TRY(e->dev_alloc((void**)&e->match_idx, (size_t)B * GT_MAX_KP * 2 * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->match_dist, (size_t)B * GT_MAX_KP * 2 * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->pairs, (size_t)B * GT_MAX_KP * 4 * sizeof(float)));
  GT_TRY(e->dev_alloc((void**)&e->npairs, (size_t)B * GT_MAX_KP * 4 * sizeof(float)));
  GT_TRY(e->dev_alloc((void**)&e->norms, (size_t)B * 8 * sizeof(float)));
  GT_TRY(e->dev_alloc((void**)&e->pair_count, (size_t)B * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->hyp_score, (size_t)B * e->cfg.ransac_max_iter * sizeof(float)));
  GT_TRY(e->dev_alloc((void**)&e->hyp_hist, (size_t)B * 256 * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->H_dev, (size_t)B * 9 * sizeof(double)));
  GT_TRY(e->dev_alloc((void**)&e->H_status, (size_t)B * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->H_stats, (size_t)B * 4 * sizeof(int)));
  GT_TRY(e->dev_alloc((void**)&e->boxes_stab_dev, (size_t)B * e->cfg.max_det * 4 * sizeof(float)));
  GT_CUDA(e, cudaMemset(e->H_status, 0, (size_t)B * sizeof(int)));
  return match_tc_build(e);
}

int match_run(gt_engine* e, const uint8_t* q, const int* nq_dev, int nq_max, const uint8_t* t, const int* nt_dev, int nt_max, int* out_idx,
              int* out_dist, int batch, size_t q_stride, size_t out_stride, cudaStream_t st) {
  (void)nt_max;
  if (e->match_mode != 0 && q == e->desc_all && t == e->desc_all + (size_t)e->cfg.max_batch * GT_MAX_KP * 32 && out_idx == e->match_idx && batch == 1)
    return match_tc_run(e, 0, 0, nq_max, e->cfg.max_batch, 0, 1, st);   // gt_match: slot 0 against the reference slot
  dim3 g((unsigned)ceil_div(nq_max, 8 * kQPW), (unsigned)batch);
  // batch > 1: queries advance by q_stride per frame with per-frame counts, the train set is shared (stride 0)
  match_kernel<<<g, 256, 0, st>>>(q, q_stride, nq_dev, batch > 1 ? 1 : 0, t, 0, nt_dev, 0, out_idx, out_dist, out_stride);
  e->launches++;
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}

int homography_run(gt_engine* e, const float* pairs, const int* counts, int B, int pair_stride, float thr, int max_iter, double* out_H,
                   int* out_status, int* out_stats, float ratio, bool full_res, const int* kp_count, cudaStream_t st) {
  ransac_prepare_kernel<<<B, 256, 0, st>>>(pairs, counts, pair_stride, (float4*)e->npairs, (Norm*)e->norms);
  GT_CUDA(e, cudaMemsetAsync(e->hyp_hist, 0, (size_t)B * 256 * sizeof(int), st));
  ransac_score_kernel<<<dim3((unsigned)ceil_div(max_iter, kScoreThreads), (unsigned)B), kScoreThreads, 0, st>>>(
      (const float4*)e->npairs, counts, pair_stride, (const Norm*)e->norms, thr, max_iter, e->cfg.seed, e->hyp_score, e->hyp_hist);
  ransac_rescore_kernel<<<B, 1024, 0, st>>>(
      (const float4*)e->npairs, counts, pair_stride, (const Norm*)e->norms, thr, max_iter, e->cfg.seed, e->hyp_score, e->hyp_hist);
  ransac_finalize_kernel<<<B, 256, 0, st>>>((const float4*)e->npairs, counts, pair_stride, (const Norm*)e->norms, e->hyp_score, thr, max_iter,
                                            e->cfg.seed, ratio, full_res ? 1 : 0, out_H, out_status, out_stats, e->cfg.max_batch,
                                            kp_count);
  e->launches += 4;
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}

int stab_match_and_fit(gt_engine* e, int B, cudaStream_t st) {
  const int R = e->cfg.max_batch;
  const size_t dstride = (size_t)GT_MAX_KP * 32, ostride = (size_t)GT_MAX_KP * 2;
  const int nkp_cur = std::min(GT_MAX_KP, e->cfg.max_features + GT_ORB_LEVELS * 64);
  const int nkp_ref = std::min(GT_MAX_KP, (int)(e->cfg.max_features * e->cfg.ref_multiplier) + GT_ORB_LEVELS * 64);
  if (e->match_mode != 0) {
    if (e->cfg.query_is_current) GT_TRY(match_tc_run(e, 0, 1, nkp_cur, R, 0, B, st));
    else GT_TRY(match_tc_run(e, R, 0, nkp_ref, 0, 1, B, st));
    e->launches -= 1;   // (the common `launches += 2` below counts build_pairs + one matcher launch)
  } else if (e->cfg.query_is_current) {
    dim3 g((unsigned)ceil_div(nkp_cur, 8 * kQPW), (unsigned)B);
    match_kernel<<<g, 256, 0, st>>>(e->desc_all, dstride, e->kp_count, 1, e->desc_all + (size_t)R * dstride, 0, e->kp_count + R, 0, e->match_idx,
                                    e->match_dist, ostride);
  } else {
    dim3 g((unsigned)ceil_div(nkp_ref, 8 * kQPW), (unsigned)B);
    match_kernel<<<g, 256, 0, st>>>(e->desc_all + (size_t)R * dstride, 0, e->kp_count + R, 0, e->desc_all, dstride, e->kp_count, 1, e->match_idx,
                                    e->match_dist, ostride);
  }
  char buf[32];
  snprintf(buf, sizeof(buf), "%.7g", (double)e->cfg.filter_ratio);  // 0.9f -> 0.9 (the reference compares in double)
  const double ratio = strtod(buf, nullptr);
  const float pt_scale = e->cfg.ransac_full_res ? 1.0f / e->cfg.downsample_ratio : 1.0f;
  build_pairs_kernel<<<B, 1024, 0, st>>>(e->match_idx, e->match_dist, e->kp_all, e->kp_count, R, e->cfg.query_is_current, ratio, pt_scale, e->pairs,
                                         e->pair_count);
  e->launches += 2;
  return homography_run(e, e->pairs, e->pair_count, B, GT_MAX_KP, e->cfg.ransac_threshold, e->cfg.ransac_max_iter, e->H_dev, e->H_status,
                        e->H_stats, e->cfg.downsample_ratio, e->cfg.ransac_full_res != 0, e->kp_count, st);
}

int warp_boxes_run(gt_engine* e, const double* H_dev, const int* status_dev, const float* in, float* out, const int* counts, int B, int stride,
                   cudaStream_t st) {
  dim3 g((unsigned)ceil_div(stride, 256), (unsigned)B);
  warp_boxes_kernel<<<g, 256, 0, st>>>(H_dev, status_dev, in, out, counts, stride);
  e->launches++;
  GT_CUDA(e, cudaGetLastError());
  return GT_OK;
}

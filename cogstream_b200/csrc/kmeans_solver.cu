// Temporal-semantic k-means on the Gram matrix: k-means++ init + Lloyd loop in ONE cooperative
// launch (reference: model/kmeans_with_time.py:41-131).
//
// Every feature distance the reference takes from torch.cdist is ||x_i||^2 + ||c_j||^2 - 2<x_i,c_j>
// with c_j a frame (init) or the mean of a set of frames M_j (Lloyd), hence
//     <x_i, c_j> = (1/n_j) * sum_{m in M_j} G[i][m]            =: S[i][j] / n_j
//     ||c_j||^2  = (1/n_j^2) * sum_{i in M_j} S[i][j]          =: q[j]
//     <c_j, c'_j> = (1/(n_j n'_j)) * sum_{i in M_j} S'[i][j]   (centre shift, :123)
// so the whole loop runs on the T x T fp64 Gram matrix (L2-resident) instead of streaming the
// [T, P*D] features 2*I + k - 1 times.  All sums are fp64 in a fixed order: deterministic, and
// unchanged memberships give a centre shift of exactly 0, as in the reference.
//
// Grid: nb co-resident CTAs (cooperative launch) separated by a software grid barrier; the init and
// the (tiny, order-sensitive) membership rebuild run on CTA 0, distances / assignment / S on all CTAs.
#include "csb_common.cuh"

namespace csb {

struct SolverArgs {
  const double* G;
  const void* ts;
  int T, k;
  int max_iter;
  double alpha, tol;
  int n_init_given, n_reseed_pool;
  int64_t* init_indices;
  const float* exp_noise;
  const int64_t* reseed_pool;
  int64_t* assignments;
  void* ct_out;
  int32_t* member_offsets;
  int32_t* members;
  csb_kmeans_status* status;
  // workspace
  float* nearest;     // [T]
  double* S[2];       // [T,k]
  double* ct[2];      // [k]
  int32_t* offs[2];   // [k+1]
  int32_t* memb[2];   // [T+k]
  unsigned* barrier;  // zeroed before launch
  int32_t* used;      // reseeds consumed so far (zeroed)
};

__device__ __forceinline__ void grid_sync(unsigned* counter, unsigned& target, unsigned nblocks) {
  __syncthreads();
  if (nblocks > 1) {
    if (threadIdx.x == 0) {
      target += nblocks;
      __threadfence();
      atomicAdd(counter, 1u);
      unsigned v;
      do {
        asm volatile("ld.acquire.gpu.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
      } while (v < target);
    }
    __syncthreads();
  }
}

template <typename TS>
__device__ __forceinline__ TS load_ts(const void* ts, int i) {
  return reinterpret_cast<const TS*>(ts)[i];
}

// ---- block-wide argmax with lowest-index tie break (scratch: 32 floats + 32 ints) ---------------
__device__ __forceinline__ int block_argmax(float v, int idx, float* sv, int* si) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  __syncthreads();
  if (lane == 0) { sv[warp] = v; si[warp] = idx; }
  __syncthreads();
  v = lane < nwarp ? sv[lane] : -1.f;
  idx = lane < nwarp ? si[lane] : 0x7fffffff;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, v, o);
    int oi = __shfl_xor_sync(0xffffffffu, idx, o);
    if (ov > v || (ov == v && oi < idx)) { v = ov; idx = oi; }
  }
  return idx;
}

// ---- k-means++ on feature distance, CTA 0 only (kmeans_with_time.py:44-62) ------------------------
// Sampling round: probs = nearest^2 / sum (:53-59), centre = argmax(probs / Exp(1) noise) -- the noise
// rows are consumed one per SAMPLING round, like the torch generator in the reference.  A round whose
// probs sum to 0 (:55-57) takes the next entry of the host-drawn random.randint pool instead.
__device__ void kmeanspp_block(const SolverArgs& a, float* sf, int* si, int* n_noise_used, int* n_degenerate) {
  const int T = a.T, k = a.k;
  const double* G = a.G;
  __shared__ int s_pick;
  for (int i = threadIdx.x; i < T; i += blockDim.x) a.nearest[i] = __int_as_float(0x7f800000);
  __syncthreads();
  int noise_row = 0, degenerate = 0;
  for (int m = 0; m < k; ++m) {
    int c;
    if (m < a.n_init_given) {
      c = (int)a.init_indices[m];
    } else {
      float part = 0.f;
      for (int i = threadIdx.x; i < T; i += blockDim.x) {
        const float nd = a.nearest[i];
        part += __fmul_rn(nd, nd);
      }
      const float sum = block_sum<float>(part, sf);
      if (sum == 0.f) {
        ++degenerate;
        if (threadIdx.x == 0) {
          int used = *a.used, r = 0;
          if (used < a.n_reseed_pool) r = (int)a.reseed_pool[used];
          else a.status->reseed_overflow = 1;
          *a.used = used + 1;
          s_pick = r;
        }
        __syncthreads();
        c = s_pick;
        __syncthreads();
      } else {
        const float* noise = a.exp_noise + (int64_t)noise_row * T;
        ++noise_row;
        float best = -1.f;
        int besti = 0x7fffffff;
        for (int i = threadIdx.x; i < T; i += blockDim.x) {
          const float nd = a.nearest[i];
          const float p = __fdiv_rn(__fmul_rn(nd, nd), sum);  // :59
          const float s = __fdiv_rn(p, noise[i]);              // multinomial: argmax(p / Exp(1))
          if (s > best) { best = s; besti = i; }
        }
        c = block_argmax(best, besti, sf, si);
      }
      if (threadIdx.x == 0) a.init_indices[m] = c;
    }
    if (m + 1 < k && a.n_init_given < k) {
      const double gcc = G[(int64_t)c * T + c];
      for (int i = threadIdx.x; i < T; i += blockDim.x) {
        double d2 = G[(int64_t)i * T + i] + gcc - 2.0 * G[(int64_t)c * T + i];
        float d = (float)sqrt(d2 > 0.0 ? d2 : 0.0);
        a.nearest[i] = fminf(a.nearest[i], d);
      }
      __syncthreads();
    }
  }
  *n_noise_used = noise_row;
  *n_degenerate = degenerate;
}

// ---- membership CSR from assignments, CTA 0 only (kmeans_with_time.py:110-120) --------------------
// A warp owns a cluster: it sweeps the assignments 32 at a time, so members come out ascending.
// Empty clusters take the next frame of the reseed pool (ascending cluster order, as :118 draws them).
template <typename TS>
__device__ void rebuild_members(const SolverArgs& a, const int64_t* assign, int32_t* offs, int32_t* memb, double* ct,
                                int* s_cnt /* smem [k+1] */) {
  const int T = a.T, k = a.k;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int j = warp; j < k; j += nwarp) {
    int cnt = 0;
    for (int i0 = 0; i0 < T; i0 += 32) {
      const int i = i0 + lane;
      const bool hit = i < T && (int)__ldcg(assign + i) == j;
      cnt += __popc(__ballot_sync(0xffffffffu, hit));
    }
    if (lane == 0) s_cnt[j] = cnt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int used = *a.used, run = 0;
    for (int j = 0; j < k; ++j) {
      const int c = s_cnt[j];
      offs[j] = run;
      if (c == 0) {
        int r = 0;
        if (used < a.n_reseed_pool) r = (int)a.reseed_pool[used];
        else a.status->reseed_overflow = 1;
        ++used;
        memb[run] = r;
        ct[j] = (double)load_ts<TS>(a.ts, r);
        run += 1;
      } else {
        run += c;
      }
    }
    offs[k] = run;
    *a.used = used;
  }
  __syncthreads();
  for (int j = warp; j < k; j += nwarp) {
    if (s_cnt[j] == 0) continue;
    int pos = offs[j];
    TS tsum = TS(0);
    for (int i0 = 0; i0 < T; i0 += 32) {
      const int i = i0 + lane;
      const bool hit = i < T && (int)__ldcg(assign + i) == j;
      const unsigned bal = __ballot_sync(0xffffffffu, hit);
      if (hit) memb[pos + __popc(bal & ((1u << lane) - 1u))] = i;
      pos += __popc(bal);
      // ascending sequential sum of member timestamps (:115)
      unsigned b = bal;
      while (b) {
        const int l = __ffs(b) - 1;
        b &= b - 1;
        tsum += load_ts<TS>(a.ts, i0 + l);
      }
    }
    if (lane == 0) ct[j] = (double)(TS)(tsum / (TS)s_cnt[j]);
  }
  __syncthreads();
}

// S[i][j] = sum_{m in M_j} G[i][m]   (all CTAs)
__device__ void compute_S(const SolverArgs& a, const int32_t* offs, const int32_t* memb, double* S) {
  const int T = a.T, k = a.k;
  const int64_t total = (int64_t)T * k;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int i = (int)(idx / k), j = (int)(idx - (int64_t)i * k);
    const int beg = __ldcg(offs + j), end = __ldcg(offs + j + 1);
    const double* g = a.G + (int64_t)i * T;
    double s = 0.0;
    for (int m = beg; m < end; ++m) s += g[__ldcg(memb + m)];
    S[idx] = s;
  }
}

template <typename TS>
__global__ void __launch_bounds__(256) kmeans_solver_kernel(SolverArgs a) {
  extern __shared__ double smem_d[];
  const int T = a.T, k = a.k;
  double* s_q[2] = {smem_d, smem_d + k};
  double* s_ct = smem_d + 2 * k;
  double* s_n = smem_d + 3 * k;
  double* s_red = smem_d + 4 * k;               // 32
  int* s_cnt = reinterpret_cast<int*>(s_red + 32);  // k+1
  __shared__ float sf[32];
  __shared__ int si[32];
  __shared__ int s_flag;

  unsigned bar = 0;
  const unsigned nb = gridDim.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  csb_kmeans_status* st = a.status;

  // ---------------- init ----------------
  if (blockIdx.x == 0) {
    if (threadIdx.x == 0) {
      st->iterations = 0; st->converged = 0; st->n_reseed_used = 0; st->degenerate_rounds = 0;
      st->reseed_overflow = 0; st->n_noise_used = 0; st->min_margin = __int_as_float(0x7f800000); st->last_shift = 0.f;
    }
    int n_noise = 0, n_deg = 0;
    if (a.n_init_given < k) kmeanspp_block(a, sf, si, &n_noise, &n_deg);
    if (threadIdx.x == 0) { st->n_noise_used = n_noise; st->degenerate_rounds = n_deg; }
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      const int c = (int)a.init_indices[j];
      a.offs[0][j] = j;
      a.memb[0][j] = c;
      a.ct[0][j] = (double)load_ts<TS>(a.ts, c);  // :66
    }
    if (threadIdx.x == 0) a.offs[0][k] = k;
  }
  grid_sync(a.barrier, bar, nb);
  compute_S(a, a.offs[0], a.memb[0], a.S[0]);
  grid_sync(a.barrier, bar, nb);

  // q for the initial centres (every CTA keeps its own copy in smem)
  int cur = 0;
  {
    const int32_t* offs = a.offs[cur];
    const int32_t* memb = a.memb[cur];
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      const int beg = __ldcg(offs + j), end = __ldcg(offs + j + 1);
      double s = 0.0;
      for (int m = beg; m < end; ++m) s += __ldcg(a.S[cur] + (int64_t)__ldcg(memb + m) * k + j);
      const double n = (double)(end - beg);
      s_q[cur][j] = s / (n * n);
      s_n[j] = n;
      s_ct[j] = __ldcg(a.ct[cur] + j);
    }
  }
  __syncthreads();

  const TS alpha = (TS)a.alpha;
  int it = 0;
  bool converged = false;
  float last_shift = 0.f;
  float my_margin = __int_as_float(0x7f800000);
  while (it < a.max_iter) {
    ++it;
    // ---------- distances + assignment (:73-104), a warp per frame ----------
    const double* S = a.S[cur];
    for (int i = blockIdx.x * nwarp + warp; i < T; i += nb * nwarp) {
      const double gii = a.G[(int64_t)i * T + i];
      const TS ti = load_ts<TS>(a.ts, i);
      float fmin_ = __int_as_float(0x7f800000), fmax_ = -fmin_;
      TS tmin_ = (TS)__int_as_float(0x7f800000), tmax_ = -tmin_;
      for (int j = lane; j < k; j += 32) {
        double d2 = gii - 2.0 * (__ldcg(S + (int64_t)i * k + j) / s_n[j]) + s_q[cur][j];
        const float d = (float)sqrt(d2 > 0.0 ? d2 : 0.0);
        TS dt = ti - (TS)s_ct[j];
        dt = dt < TS(0) ? -dt : dt;
        fmin_ = fminf(fmin_, d); fmax_ = fmaxf(fmax_, d);
        tmin_ = dt < tmin_ ? dt : tmin_; tmax_ = dt > tmax_ ? dt : tmax_;
      }
      fmin_ = warp_min(fmin_); fmax_ = warp_max(fmax_);
      tmin_ = warp_min(tmin_); tmax_ = warp_max(tmax_);
      const float fspread = __fsub_rn(fmax_, fmin_);
      const TS tspread = tmax_ - tmin_;
      float best = __int_as_float(0x7f800000), second = best;
      int bestj = 0x7fffffff;
      for (int j = lane; j < k; j += 32) {
        double d2 = gii - 2.0 * (__ldcg(S + (int64_t)i * k + j) / s_n[j]) + s_q[cur][j];
        const float d = (float)sqrt(d2 > 0.0 ? d2 : 0.0);
        TS dt = ti - (TS)s_ct[j];
        dt = dt < TS(0) ? -dt : dt;
        const float nf = (fmax_ > fmin_) ? __fdiv_rn(__fsub_rn(d, fmin_), fspread) : 0.f;  // :86-91
        const TS nt = (tmax_ > tmin_) ? (dt - tmin_) / tspread : TS(0);                        // :94-99
        float fin;                                                                               // :101
        if (sizeof(TS) == 8) {
          const double v = __dadd_rn((double)__fmul_rn(nf, nf), __dmul_rn((double)alpha, __dmul_rn((double)nt, (double)nt)));
          fin = __double2float_rn(sqrt(v));
        } else {
          const float ntf = (float)nt;
          fin = __fsqrt_rn(__fadd_rn(__fmul_rn(nf, nf), __fmul_rn((float)alpha, __fmul_rn(ntf, ntf))));
        }
        if (fin < best) { second = best; best = fin; bestj = j; }
        else if (fin < second) second = fin;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const float os = __shfl_xor_sync(0xffffffffu, second, o);
        const int oj = __shfl_xor_sync(0xffffffffu, bestj, o);
        const float hi = fmaxf(best, ob);
        second = fminf(hi, fminf(second, os));
        if (ob < best || (ob == best && oj < bestj)) { best = ob; bestj = oj; }
      }
      if (lane == 0) {
        a.assignments[i] = bestj;  // :104 lowest index on ties
        if (k > 1) my_margin = fminf(my_margin, second - best);
      }
    }
    grid_sync(a.barrier, bar, nb);

    // ---------- new centres: membership, times, re-seeds (CTA 0) ----------
    const int nxt = cur ^ 1;
    if (blockIdx.x == 0) rebuild_members<TS>(a, a.assignments, a.offs[nxt], a.memb[nxt], a.ct[nxt], s_cnt);
    grid_sync(a.barrier, bar, nb);
    compute_S(a, a.offs[nxt], a.memb[nxt], a.S[nxt]);
    grid_sync(a.barrier, bar, nb);

    // ---------- centre shift (:123-125); every CTA evaluates it identically ----------
    double part = 0.0;
    {
      const int32_t* offs = a.offs[nxt];
      const int32_t* memb = a.memb[nxt];
      for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const int beg = __ldcg(offs + j), end = __ldcg(offs + j + 1);
        double sn = 0.0, sc = 0.0;
        for (int m = beg; m < end; ++m) {
          const int64_t row = (int64_t)__ldcg(memb + m) * k + j;
          sn += __ldcg(a.S[nxt] + row);
          sc += __ldcg(a.S[cur] + row);
        }
        const double n = (double)(end - beg);
        const double qn = sn / (n * n);
        const double cross = sc / (n * s_n[j]);
        double sh2 = (qn + s_q[cur][j]) - 2.0 * cross;
        part += sqrt(sh2 > 0.0 ? sh2 : 0.0);
        s_q[nxt][j] = qn;
      }
    }
    const double shift_feat = block_sum<double>(part, s_red);
    double tpart = 0.0;
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      const double ctn = __ldcg(a.ct[nxt] + j);
      const TS d = (TS)ctn - (TS)s_ct[j];
      tpart += (double)(d * d);
    }
    const double shift_time = sqrt(block_sum<double>(tpart, s_red));
    __syncthreads();
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      s_n[j] = (double)(__ldcg(a.offs[nxt] + j + 1) - __ldcg(a.offs[nxt] + j));
      s_ct[j] = __ldcg(a.ct[nxt] + j);
    }
    __syncthreads();
    cur = nxt;
    const double shift = shift_feat + shift_time;
    last_shift = (float)shift;
    if (shift <= a.tol) { converged = true; break; }  // :130
  }

  // ---------------- outputs ----------------
  if (lane == 0 && my_margin < __int_as_float(0x7f800000))
    atomicMin(reinterpret_cast<unsigned*>(&st->min_margin), __float_as_uint(fmaxf(my_margin, 0.f)));
  if (blockIdx.x == 0) {
    for (int j = threadIdx.x; j <= k; j += blockDim.x) a.member_offsets[j] = __ldcg(a.offs[cur] + j);
    const int nmem = __ldcg(a.offs[cur] + k);
    for (int m = threadIdx.x; m < nmem; m += blockDim.x) a.members[m] = __ldcg(a.memb[cur] + m);
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
      if (sizeof(TS) == 8) reinterpret_cast<double*>(a.ct_out)[j] = s_ct[j];
      else reinterpret_cast<float*>(a.ct_out)[j] = (float)s_ct[j];
    }
    if (threadIdx.x == 0) {
      st->iterations = it;
      st->converged = converged ? 1 : 0;
      st->n_reseed_used = *a.used;
      st->last_shift = last_shift;
    }
  }
}

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace csb

using namespace csb;

extern "C" {

size_t csb_kmeans_workspace_bytes(int64_t T, int64_t k) {
  if (T <= 0 || k <= 0) return 0;
  size_t b = 0;
  b += align_up(sizeof(float) * T, 256);               // nearest
  b += 2 * align_up(sizeof(double) * T * k, 256);      // S
  b += 2 * align_up(sizeof(double) * k, 256);          // ct
  b += 2 * align_up(sizeof(int32_t) * (k + 1), 256);   // offs
  b += 2 * align_up(sizeof(int32_t) * (T + k), 256);   // memb
  b += 256;                                            // barrier + used
  return b;
}

int csb_kmeans_solve(const double* G, const void* timestamps, const csb_kmeans_params* p, int64_t* init_indices,
                     const float* exp_noise, const int64_t* reseed_pool, int64_t* assignments, void* centroid_times,
                     int32_t* member_offsets, int32_t* members, csb_kmeans_status* status, void* workspace,
                     size_t workspace_bytes, csb_stream_t stream) {
  CSB_DEVICE_OR_RETURN();
  CSB_REQUIRE(p, "kmeans_solve: null params");
  CSB_REQUIRE(G && timestamps && init_indices && assignments && centroid_times && member_offsets && members && status && workspace,
              "kmeans_solve: null pointer");
  CSB_REQUIRE(p->k >= 1 && p->T > p->k, "kmeans_solve: need T > k >= 1 (T=%lld k=%lld); T <= k is the caller's early return",
              (long long)p->T, (long long)p->k);
  CSB_REQUIRE(p->T < (1ll << 30) && p->k <= 4096, "kmeans_solve: T or k too large (k <= 4096)");
  CSB_REQUIRE(p->ts_dtype == CSB_F32 || p->ts_dtype == CSB_F64, "kmeans_solve: timestamps must be f32 or f64");
  CSB_REQUIRE(p->n_init_given >= 1 && p->n_init_given <= p->k, "kmeans_solve: n_init_given must be in [1,k]");
  CSB_REQUIRE(p->n_init_given == p->k || exp_noise, "kmeans_solve: exp_noise required when centres are sampled");
  CSB_REQUIRE(p->n_reseed_pool == 0 || reseed_pool, "kmeans_solve: reseed_pool is null");
  CSB_REQUIRE(p->max_iteration >= 0, "kmeans_solve: max_iteration < 0");
  if (workspace_bytes < csb_kmeans_workspace_bytes(p->T, p->k)) {
    set_error("kmeans_solve: workspace %zu < %zu", workspace_bytes, csb_kmeans_workspace_bytes(p->T, p->k));
    return CSB_ERR_WORKSPACE;
  }
  const int64_t T = p->T, k = p->k;
  SolverArgs a;
  a.G = G; a.ts = timestamps; a.T = (int)T; a.k = (int)k; a.max_iter = p->max_iteration; a.alpha = p->alpha; a.tol = p->tol;
  a.n_init_given = (int)p->n_init_given; a.n_reseed_pool = (int)p->n_reseed_pool;
  a.init_indices = init_indices; a.exp_noise = exp_noise; a.reseed_pool = reseed_pool; a.assignments = assignments;
  a.ct_out = centroid_times; a.member_offsets = member_offsets; a.members = members; a.status = status;
  char* w = reinterpret_cast<char*>(workspace);
  a.nearest = reinterpret_cast<float*>(w); w += align_up(sizeof(float) * T, 256);
  for (int b = 0; b < 2; ++b) { a.S[b] = reinterpret_cast<double*>(w); w += align_up(sizeof(double) * T * k, 256); }
  for (int b = 0; b < 2; ++b) { a.ct[b] = reinterpret_cast<double*>(w); w += align_up(sizeof(double) * k, 256); }
  for (int b = 0; b < 2; ++b) { a.offs[b] = reinterpret_cast<int32_t*>(w); w += align_up(sizeof(int32_t) * (k + 1), 256); }
  for (int b = 0; b < 2; ++b) { a.memb[b] = reinterpret_cast<int32_t*>(w); w += align_up(sizeof(int32_t) * (T + k), 256); }
  a.barrier = reinterpret_cast<unsigned*>(w);
  a.used = reinterpret_cast<int32_t*>(w + 128);
  cudaStream_t s = as_stream(stream);
  cudaError_t e = cudaMemsetAsync(w, 0, 256, s);
  if (e != cudaSuccess) { set_error("kmeans_solve: memset: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }

  const size_t smem = sizeof(double) * (4 * k + 32) + sizeof(int) * (k + 1) + 16;
  auto kern = (p->ts_dtype == CSB_F64) ? kmeans_solver_kernel<double> : kmeans_solver_kernel<float>;
  if (smem > 48 * 1024) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("kmeans_solve: smem attr: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
  }
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 256, smem);
  if (e != cudaSuccess || per_sm < 1) { set_error("kmeans_solve: occupancy query failed"); return CSB_ERR_CUDA; }
  int dev = 0, sms = kNumSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  int64_t nb = (T * k + 2047) / 2048;
  if (nb > sms) nb = sms;
  if (nb < 1) nb = 1;
  void* args[] = {&a};
  e = cudaLaunchCooperativeKernel((const void*)kern, dim3((unsigned)nb), dim3(256), args, smem, s);
  if (e != cudaSuccess) { set_error("kmeans_solve: launch: %s", cudaGetErrorString(e)); return CSB_ERR_CUDA; }
  count_launch(1);
  return CSB_OK;
}

}  // extern "C"
